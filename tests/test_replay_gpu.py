"""Replay exchange on the device: mopa_rollout_pack / mopa_replay_append against numpy mirrors, queueing of bursts, the
overflow error, and - on a box with two GPUs - the NCCL exchange with the 1-GPU-vs-2-GPU invariance of the collected
transition multiset."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sorted_rows(a):
    return a[np.lexsort(a.T[::-1])]


def test_pack_and_append_move_every_record_once_in_order(push_model):
    import torch

    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.replay import ReplicatedReplay
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner

    n = 64
    venv = VecSawyerPushObstacle(n, seed=21, max_episode_steps=30)
    runner = NativeMoPARolloutRunner(venv, MoPAConfig(max_iter=100, reuse_data=True, max_reuse_data=15), policy=CounterPolicy(torch, venv.dev, 2))
    rep = ReplicatedReplay(torch, venv.dev, capacity=1 << 14, slab_capacity=8)   # far below the burst size: records queue up
    seen = 0
    for t in range(40):
        runner.tick()
        rep.exchange(runner)
        blk = rep.last_block
        k = int(blk[0, :1].view(torch.int32)[0])
        assert 0 <= k <= 8 and torch.all(blk[0, 1:] == 0)
        seen += k
    total = runner.counters["transitions"]
    assert total > 8 * 40 and seen < total           # the bursts did not fit: some records are still queued
    for _ in range((total - seen) // 8 + 2):        # drain the queue (no ticks: nothing new is emitted)
        rep.exchange(runner)
    torch.cuda.synchronize()
    assert rep.device_size() == total == runner.counters["transitions"]
    assert torch.equal(rep.ring[:total], runner.transitions[:total])          # world = 1: the replica is the local ring, in order
    assert rep.sample(16).shape == (16, 92)


def test_append_is_rank_major_and_wraps():
    import torch

    from mopa_rl_b200.capi import check, lib

    L = lib()
    L.mopa_replay_append.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]
    world, cap, ring_cap = 3, 5, 16
    rng = np.random.default_rng(0)
    ring = torch.zeros(ring_cap, 92, device="cuda")
    size2 = torch.zeros(2, dtype=torch.int64, device="cuda")
    mirror, msize, parity = np.zeros((ring_cap, 92), np.float32), 0, 0
    for step in range(6):
        blocks = np.zeros((world, 1 + cap, 92), np.float32)
        for rk in range(world):
            k = int(rng.integers(0, cap + 1))
            blocks[rk, 0, :1] = np.array([k], np.int32).view(np.float32)
            blocks[rk, 1:1 + k] = rng.random((k, 92)).astype(np.float32)
            blocks[rk, 1 + k:] = -7.0                                           # stale rows beyond the count must be ignored
            for i in range(k):
                mirror[(msize + i) % ring_cap] = blocks[rk, 1 + i]
            msize += k
        d = torch.as_tensor(blocks.reshape(-1, 92), device="cuda")
        check(L.mopa_replay_append(d.data_ptr(), world, cap, ring.data_ptr(), ring_cap, size2.data_ptr(), parity, None))
        parity ^= 1
        torch.cuda.synchronize()
        assert int(size2[parity]) == msize
        assert np.array_equal(ring.cpu().numpy(), mirror)
    assert msize > ring_cap                                                     # the ring wrapped


def test_queue_overflow_raises(push_model):
    import torch

    from mopa_rl_b200.capi import MopaError
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.replay import ReplicatedReplay
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner

    venv = VecSawyerPushObstacle(64, seed=5, max_episode_steps=30)
    runner = NativeMoPARolloutRunner(venv, MoPAConfig(max_iter=50, reuse_data=True), policy=CounterPolicy(torch, venv.dev, 2), transition_capacity=64)
    rep = ReplicatedReplay(torch, venv.dev, capacity=1 << 12, slab_capacity=1)   # one record per tick leaves, dozens arrive
    with pytest.raises(MopaError, match="overwritten"):
        for t in range(60):
            runner.tick()
            rep.exchange(runner)
            torch.cuda.synchronize()


def _nccl_worker(rank, world, port, n_per, ticks, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist

    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.replay import ReplicatedReplay
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    cfg = MoPAConfig(max_iter=100, reuse_data=True, max_reuse_data=15, seed=9)
    venv = VecSawyerPushObstacle(n_per, seed=33, device=rank, env_id_offset=rank * n_per, max_episode_steps=30)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, dev, 4))
    rep = ReplicatedReplay(torch, dev, capacity=1 << 15, slab_capacity=32)
    for t in range(ticks):
        runner.tick(wait_rrt=True)            # deterministic timeline: every RRT batch is adopted one tick after its launch
        rep.exchange(runner)
    for _ in range(64):                        # drain the queues (same number of collective calls on every rank)
        rep.exchange(runner)
    torch.cuda.synchronize()
    size = rep.device_size()
    np.save(os.path.join(out_dir, "ring%d.npy" % rank), rep.ring[:size].cpu().numpy())
    np.save(os.path.join(out_dir, "local%d.npy" % rank), runner.transitions[:runner.counters["transitions"]].cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_exchange_and_gpu_count_invariance(tmp_path):
    """Two ranks (one GPU each) collect from 2 x 48 envs and all-gather their records; one rank with 96 envs collects
    the same global env ids.  Every replica holds the same ring, and the multiset of records does not depend on the
    GPU count (RNG streams are keyed by global env id, planner problems by (env id, plan #))."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    n_per, ticks = 48, 45
    port = 23456 + (os.getpid() % 1000)
    mp.spawn(_nccl_worker, args=(2, port, n_per, ticks, str(tmp_path)), nprocs=2, join=True)
    ring0, ring1 = np.load(tmp_path / "ring0.npy"), np.load(tmp_path / "ring1.npy")
    local = np.concatenate([np.load(tmp_path / "local0.npy"), np.load(tmp_path / "local1.npy")])
    assert len(ring0) == len(local) > 4 * n_per
    assert np.array_equal(ring0, ring1), "replicas differ"
    assert np.array_equal(_sorted_rows(ring0), _sorted_rows(local)), "the replicated ring is not the union of the ranks' records"

    # the same 96 global envs on one GPU
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner

    cfg = MoPAConfig(max_iter=100, reuse_data=True, max_reuse_data=15, seed=9)
    venv = VecSawyerPushObstacle(2 * n_per, seed=33, max_episode_steps=30)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 4))
    for t in range(ticks):
        runner.tick(wait_rrt=True)
    torch.cuda.synchronize()
    one = runner.transitions[:runner.counters["transitions"]].cpu().numpy()
    assert len(one) == len(ring0)
    assert np.array_equal(_sorted_rows(one), _sorted_rows(ring0)), "1-GPU and 2-GPU runs collected different transitions"
