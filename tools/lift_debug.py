"""Per-step / per-env comparison of the lift env kernel with its oracle (debug aid for tests/test_env_gpu.py)."""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_env_gpu as T
from mopa_rl_b200.dynmodel import DynModel
from mopa_rl_b200.envs import VecSawyerLiftObstacle, lift_reset_state
from mopa_rl_b200.model import load_model
from oracle.env_oracle import LiftEnvOracle

model = load_model("SawyerLiftObstacle-v0")
n = 16
venv = VecSawyerLiftObstacle(n, seed=13, max_episode_steps=70)
venv.reset()
dm = DynModel(model)
q0, v0 = lift_reset_state(model, 13, np.arange(n), np.zeros(n, dtype=np.int64))
j1 = model.get_joint_qpos_addr("right_j1")
q0[1::4, j1] = -0.6
venv.set_state(np.arange(n), q0, v0)
envs = [LiftEnvOracle(model, dm, max_episode_steps=70) for _ in range(n)]
for i, e in enumerate(envs): e.reset_to(q0[i], v0[i])
rng = np.random.default_rng(4)
a = model.get_joint_qpos_addr("cube")[0]; va = model.get_joint_qvel_addr("cube")[0]
def step(act, isp, tag):
    venv.step(torch.as_tensor(act, device="cuda"), torch.as_tensor(isp, device="cuda")); torch.cuda.synchronize()
    gq, gv, grew, gn = venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy(), venv.reward.cpu().numpy(), venv.ncon.cpu().numpy()
    for i, e in enumerate(envs):
        ob, r, d = e.step(act[i].astype(np.float64), bool(isp[i]))
        dq = np.abs(gq[i] - e.qpos); dv = np.abs(gv[i] - e.qvel)
        print(tag, i, "dq %.2e @%d dv %.2e @%d rew %.4f/%.4f ncon %d/%d" % (dq.max(), dq.argmax(), dv.max(), dv.argmax(), grew[i], r, gn[i], e.ncon), e.contacts)
for s in range(2):
    act = np.zeros((n, 8), np.float32); act[:, :7] = rng.uniform(-0.2, 0.2, (n, 7)); act[:, 7] = -1.0
    step(act, np.zeros(n, np.uint8), "open%d" % s)
q, v = np.stack([e.qpos for e in envs]), np.stack([e.qvel for e in envs])
for i, e in enumerate(envs):
    if i % 2 == 0 or i % 4 == 1:
        mid, quat = T._lift_fingertip_frame(model, dm, e)
        q[i, a:a + 3], q[i, a + 3:a + 7], v[i, va:va + 6] = mid, quat, 0.0
    e.set_state(q[i], v[i]); e.prev_state = None
venv.set_state(np.arange(n), q, v); venv.reset_prev_state()
for s in range(5):
    act = rng.uniform(-1, 1, (n, 8)).astype(np.float32) if s >= 2 else np.zeros((n, 8), np.float32)
    act[:, 7] = 0.004
    isp = np.zeros(n, np.uint8)
    if s >= 3: isp[::2] = 1; act[::2, :7] *= 0.08
    step(act, isp, "grasp%d" % s)
