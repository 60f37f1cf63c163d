"""Per-environment Newton step totals of one env.step (diagnostics, MOPA_ENV_PROF=2)."""
import os, sys
os.environ["MOPA_ENV_PROF"] = "2"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from mopa_rl_b200.envs import VecSawyerPushObstacle
n = 4096
venv = VecSawyerPushObstacle(n, seed=1234)
venv.reset()
tot = []
for k in range(6):
    a = torch.rand(n, 8, device="cuda") * 2 - 1
    venv.step(a)
    torch.cuda.synchronize()
    v = venv.ncon.cpu().numpy()
    it, nc = v // 1000, v % 1000
    tot.append(it)
    print("step %d: newton steps per env.step: mean %.1f  p10 %d  p50 %d  p90 %d  p99 %d  max %d; ncon mean %.2f; corr(iters, ncon) %.2f" % (
        k, it.mean(), *np.percentile(it, [10, 50, 90, 99]).astype(int), it.max(), nc.mean(), np.corrcoef(it, nc)[0, 1]))
tot = np.array(tot)
print("corr between consecutive env.steps of the same env:", np.mean([np.corrcoef(tot[k], tot[k + 1])[0, 1] for k in range(2, 5)]))
for lo, hi in ((0, 76), (76, 90), (90, 120), (120, 200), (200, 10000)):
    print("  envs with %d <= steps < %d: %.1f%%" % (lo, hi, 100 * np.mean((tot[-1] >= lo) & (tot[-1] < hi))))
