"""Long-run behaviour of the native runner: throughput and RRT backlog per window of ticks."""
import sys, time; sys.path.insert(0, '.')
import torch
from mopa_rl_b200.envs import VecSawyerPushObstacle
from mopa_rl_b200.replay import ReplicatedReplay
from mopa_rl_b200.rollout import MoPAConfig, NativeMoPARolloutRunner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 300
venv = VecSawyerPushObstacle(n, seed=1234)
r = NativeMoPARolloutRunner(venv, MoPAConfig(max_iter=1000, reuse_data=True, max_reuse_data=15))   # the push preset, as bench.py runs it
rep = ReplicatedReplay(torch, venv.dev, capacity=1 << 20, slab_capacity=max(1024, n))
prev = r.counters; torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(1, ticks + 1):
    r.tick()
    rep.exchange(r)
    if k % 25 == 0:
        c = r.counters; t1 = time.perf_counter()
        print("ticks %4d  %.1f ms/tick  env-steps/s %7.0f  waiting %4d  rrt queued %5d done %5d  episodes %d  records %d (relabelled %d) queued for exchange %d" % (
            k, (t1 - t0) / 25 * 1e3, (c["env_steps"] - prev["env_steps"]) / (t1 - t0), c["waiting"], c["rrt_problems"], c["mp"] + c["approximate"], c["episodes"],
            c["transitions"], c["reused"], c["transitions"] - rep.device_size()), " rrt batches: last %.1f ms, n %d, mean %.1f ms, %.1f ticks" % r.rrt_stats())
        prev, t0 = c, time.perf_counter()
print(r.counters)
