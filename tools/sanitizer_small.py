import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import planner_setup, random_qpos
from mopa_rl_b200.capi import NativePlanner
from mopa_rl_b200.model import load_model
mode = sys.argv[1]
if mode == "validity":
    m = load_model("SawyerPushObstacle-v0")
    ign, passive, ref = planner_setup(m)
    pl = NativePlanner(m, passive, ign, -0.002, 0.1, seed=1)
    for n in (3000, 67000):
        q = random_qpos(m, n, 5, ref)
        v = pl.is_valid_host(q, flags=1, return_words=True)[1]
        print(n, "valid fraction", float((v & 1).mean()))
else:
    from mopa_rl_b200.envs import VecSawyerPushObstacle, VecSawyerLiftObstacle
    for cls, n in ((VecSawyerPushObstacle, 30), (VecSawyerLiftObstacle, 12)):
        venv = cls(n, seed=3); venv.reset()
        a = torch.rand(n, 8, device='cuda') * 2 - 1
        venv.step(a); torch.cuda.synchronize()
        print(cls.__name__, "ok", float(venv.reward.sum()))
