"""Host-side logic of the multi-GPU transition exchange (mopa_rl_b200/replay.py), exercised with world_size-2 gloo on CPU.

The product path packs and appends with CUDA kernels (mopa_rollout_pack / mopa_replay_append); here a test double
replaces exactly those two device calls (and the stream bookkeeping) with numpy-level mirrors, so that the protocol
around them - double-buffered blocks, header rows, the all-gather, rank-major append, ping-pong size counter, queueing
of bursts - runs unchanged under gloo.  The kernels themselves are compared with the same mirrors in
tests/test_replay_gpu.py."""
import contextlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeRunner:
    """Stands in for the native runner: a FIFO of emitted records; pack() moves at most `capacity` of them."""

    def __init__(self):
        self.queue = []

    def emit(self, rows):
        self.queue.extend(list(rows))

    def pack(self, send, capacity):
        k = min(len(self.queue), capacity)
        send.zero_()
        send[0, 0] = torch.tensor([k], dtype=torch.int32).view(torch.float32)[0]
        if k:
            send[1:1 + k] = torch.stack(self.queue[:k])
        del self.queue[:k]


def host_replay_class():
    from mopa_rl_b200.replay import ReplicatedReplay

    class HostReplay(ReplicatedReplay):
        def _init_streams(self, overlap):
            pass

        def _wait_block(self, b):
            pass

        def _side_stream(self):
            return contextlib.nullcontext()

        def _mark_done(self, b):
            pass

        def _append(self, blocks):   # mirror of replay_append_kernel
            rows = 1 + self.slab_capacity
            size0 = int(self._size2[self._parity])
            off = 0
            for rk in range(self.world):
                blk = blocks[rk * rows:(rk + 1) * rows]
                k = int(blk[0, :1].view(torch.int32)[0])
                k = max(0, min(k, self.slab_capacity))
                for i in range(k):
                    self.ring[(size0 + off + i) % self.capacity] = blk[1 + i]
                off += k
            self._size2[1 - self._parity] = size0 + off

    return HostReplay


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rep = host_replay_class()(torch, torch.device("cpu"), capacity=64, slab_capacity=4)
    runner = FakeRunner()
    rng = np.random.default_rng(rank)
    sent = []
    for tick in range(9):
        k = [3, 0, 7, 0, 0][(tick + rank) % 5] if tick < 6 else 0     # ragged, sometimes empty, one burst above the block capacity
        if k:
            rec = torch.as_tensor(rng.random((k, 92)).astype(np.float32))
            rec[:, 51] = rank
            sent.append(rec)
            runner.emit(rec)
        rep.exchange(runner)
    size = rep.device_size()
    out[rank] = (rep.ring[: min(size, 64)].clone().numpy(), size, torch.cat(sent).numpy(), len(runner.queue), rep.bytes_exchanged)
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_gloo_world2():
    """Identical rings on both ranks, every record exactly once, bursts above the block capacity drain over later ticks."""
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ring0, size0, sent0, left0, bytes0 = out[0]
    ring1, size1, sent1, left1, _ = out[1]
    assert left0 == left1 == 0
    assert size0 == size1 == len(sent0) + len(sent1)
    assert np.array_equal(ring0, ring1), "replicas must hold identical rings"
    allsent = np.concatenate([sent0, sent1])
    assert sorted(map(bytes, allsent)) == sorted(map(bytes, ring0[:size0]))
    assert bytes0 == 9 * 2 * 5 * 92 * 4   # 9 ticks x 2 ranks x (1 + 4) rows: fixed-size blocks, no per-environment padding
    # rank-major within a tick: the first tick holds rank 0's three records, then rank 1's (none: k = 0 for rank 1 at tick 0)
    assert np.array_equal(ring0[:3], sent0[:3])


def test_single_process_wraps_and_keeps_order():
    rep = host_replay_class()(torch, torch.device("cpu"), capacity=8, slab_capacity=4)
    runner = FakeRunner()
    for i in range(5):
        runner.emit(torch.full((3, 92), float(i)))
        rep.exchange(runner)
    assert rep.device_size() == 15 and not runner.queue
    assert set(rep.ring[:, 0].tolist()) <= {2.0, 3.0, 4.0}
    assert rep.sample(6).shape == (6, 92)
    # a burst larger than the block: FIFO order is kept across ticks
    rep2 = host_replay_class()(torch, torch.device("cpu"), capacity=32, slab_capacity=4)
    r2 = FakeRunner()
    r2.emit(torch.arange(10, dtype=torch.float32)[:, None].repeat(1, 92))
    for _ in range(3):
        rep2.exchange(r2)
    assert rep2.device_size() == 10 and rep2.ring[:10, 0].tolist() == list(map(float, range(10)))


def test_product_class_refuses_cpu():
    import pytest

    from mopa_rl_b200.replay import ReplicatedReplay

    with pytest.raises(RuntimeError):
        ReplicatedReplay(torch, torch.device("cpu"), capacity=8, slab_capacity=4)
