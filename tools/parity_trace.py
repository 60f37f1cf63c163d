"""Per-env.step error trace of one environment of the bench configuration against the scalar loop:
tools/parity_trace.py <global env id> [horizon]"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch  # noqa: E402

from mopa_rl_b200.envs import VecSawyerPushObstacle  # noqa: E402
from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner  # noqa: E402
from test_scale_parity_gpu import BENCH_SEED_CFG, BENCH_SEED_ENV, BENCH_SEED_POLICY, _scalar_episode  # noqa: E402

gid = int(sys.argv[1])
horizon = int(sys.argv[2]) if len(sys.argv) > 2 else 250
cfg = MoPAConfig(max_iter=1000, reuse_data=True, max_reuse_data=15, seed=BENCH_SEED_CFG)
venv = VecSawyerPushObstacle(4, seed=BENCH_SEED_ENV, max_episode_steps=horizon, env_id_offset=gid)
runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, BENCH_SEED_POLICY))
log, prev, done = {}, 0, False
for t in range(horizon * 2):
    runner.tick(wait_rrt=True)
    L = int(venv.ep_len[0])
    if L < prev:
        break
    prev = L
    if L > 0:
        log[L] = (venv.qpos[0].cpu().numpy(), venv.qvel[0].cpu().numpy(), int(venv.ncon[0]), int(venv.work[0]))
_, sq, sv, srec = _scalar_episode((gid, horizon))
print("step  |dq|      |dv|      ncon newton-steps")
for k in sorted(log):
    if k <= len(sq):
        eq, ev = np.abs(log[k][0] - sq[k - 1]).max(), np.abs(log[k][1] - sv[k - 1]).max()
        if k < 5 or k % 10 == 0 or eq > 1e-7:
            print("%4d  %.2e  %.2e  %d  %d  argmax dof %d" % (k, eq, ev, log[k][2], log[k][3], int(np.abs(log[k][1] - sv[k - 1]).argmax())))
