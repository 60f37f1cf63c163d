"""Host-side logic of the multi-GPU transition exchange, exercised with world_size-2 gloo on CPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mopa_rl_b200.replay import ReplicatedReplay

    rep = ReplicatedReplay(torch, torch.device("cpu"), capacity=64, slab_capacity=16)
    rng = np.random.default_rng(rank)
    sent = []
    for tick in range(7):
        k = [3, 0, 5][(tick + rank) % 3]                     # ragged, sometimes empty
        rec = None
        if k:
            rec = torch.as_tensor(rng.random((k, 92)).astype(np.float32))
            rec[:, 51] = rank
            sent.append(rec)
        rep.exchange(rec)
    out[rank] = (rep.ring[: min(rep.size, 64)].clone().numpy(), rep.size, torch.cat(sent).numpy())
    dist.barrier()
    dist.destroy_process_group()


def _slab_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mopa_rl_b200.replay import ReplicatedReplay

    n = 6   # env rows per rank
    rep = ReplicatedReplay(torch, torch.device("cpu"), capacity=64)
    rng = np.random.default_rng(10 + rank)
    sent = []
    for tick in range(6):
        slab = torch.as_tensor(rng.random((n, 92)).astype(np.float32))
        slab[:, 51] = rank
        flags = torch.as_tensor(((np.arange(n) + tick + rank) % 3 == 0).astype(np.uint8))   # ragged, sometimes empty
        if tick == 4:
            flags[:] = 0
        sent.append(slab[flags.bool()])
        rep.exchange_slab(slab, flags)
    size = rep.device_size()
    out[rank] = (rep.ring[: min(size, 64)].clone().numpy(), size, torch.cat(sent).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_slab_exchange_gloo_world2():
    """exchange_slab (the native runner's fixed-shape, sync-free path): identical rings, every record exactly once."""
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_slab_worker, args=(2, port, out), nprocs=2, join=True)
    ring0, size0, sent0 = out[0]
    ring1, size1, sent1 = out[1]
    assert size0 == size1 == len(sent0) + len(sent1)
    assert np.array_equal(ring0, ring1)
    allsent = np.concatenate([sent0, sent1])
    assert sorted(map(bytes, allsent)) == sorted(map(bytes, ring0[:size0]))


def test_slab_exchange_single_process_wraps():
    from mopa_rl_b200.replay import ReplicatedReplay

    rep = ReplicatedReplay(torch, torch.device("cpu"), capacity=8)
    for i in range(5):
        rep.exchange_slab(torch.full((4, 92), float(i)), torch.tensor([1, 0, 1, 1], dtype=torch.uint8))
    assert rep.device_size() == 15 and set(rep.ring[:, 0].tolist()) <= {2.0, 3.0, 4.0}
    assert rep.sample(6).shape == (6, 92)


def test_replicated_replay_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ring0, size0, sent0 = out[0]
    ring1, size1, sent1 = out[1]
    assert size0 == size1 == len(sent0) + len(sent1)
    assert np.array_equal(ring0, ring1), "replicas must hold identical rings"
    # every record each rank sent is present exactly once
    allsent = np.concatenate([sent0, sent1])
    assert sorted(map(bytes, allsent)) == sorted(map(bytes, ring0[:size0]))


def test_single_process_replay_wraps():
    from mopa_rl_b200.replay import ReplicatedReplay, pack_counts_and_slab

    rep = ReplicatedReplay(torch, torch.device("cpu"), capacity=8, slab_capacity=4)
    for i in range(5):
        rep.exchange(torch.full((3, 92), float(i)))
    assert rep.size == 15 and set(rep.ring[:, 0].tolist()) <= {2.0, 3.0, 4.0}
    rep.exchange(None)
    assert rep.size == 15
    c, slab = pack_counts_and_slab(torch, None, 4, torch.device("cpu"))
    assert int(c) == 0 and slab.shape == (4, 92)
    with pytest.raises(ValueError):
        pack_counts_and_slab(torch, torch.zeros(5, 92), 4, torch.device("cpu"))
    assert rep.sample(6).shape == (6, 92)
