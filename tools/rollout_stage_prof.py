"""Stage clocks of the env-step kernel inside the real rollout (grouped launch order).  Needs MOPA_ENV_PROF=1."""
import ctypes as C, os, sys
os.environ["MOPA_ENV_PROF"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from mopa_rl_b200.envs import VecSawyerPushObstacle
from mopa_rl_b200.rollout import MoPAConfig, NativeMoPARolloutRunner
venv = VecSawyerPushObstacle(4096, seed=1234)
r = NativeMoPARolloutRunner(venv, MoPAConfig(max_iter=1000, reuse_data=True))
for _ in range(60):
    r.tick()
torch.cuda.synchronize()
out = (C.c_uint64 * 32)()
venv._L.mopa_env_debug_prof.argtypes = [C.c_void_p, C.c_void_p]
venv._L.mopa_env_debug_prof(venv.h, out)
v = np.array(list(out), dtype=np.float64)
names = ["kinematics", "inertia+RNE", "CRBA+forces", "chol+qacc0", "contacts", "rows+solver", "integrate"]
tot = v[2:16].sum() + v[23:27].sum()
for k in range(1, 8):
    print("stage %d %-12s work %5.1f%%  wait %5.1f%%" % (k, names[k - 1], 100 * v[2 * k] / tot, 100 * v[2 * k + 1] / tot))
print("within 5: broadphase %.1f%%  narrowphase %.1f%%;  within 6: rows %.1f%%  Newton %.1f%%" % tuple(100 * v[k] / tot for k in (23, 24, 25, 26)))
print("newton steps per substep %.3f" % (v[20] / max(v[22], 1)))
print("clocks per constrained substep (lane 0 of every warp, all stages): %.0f; constrained substeps %.3g" % (tot / max(v[22], 1), v[22]))
