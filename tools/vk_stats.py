"""Per pair class: (query, pair) items tested / surviving the cull / offending / MPR items that reach the portal refinement.
Needs a diagnostics build: MOPA_EXTRA_NVCC=-DMOPA_VK_STATS python -m mopa_rl_b200.build --force"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from helpers import planner_setup, random_qpos  # noqa: E402
from mopa_rl_b200.capi import NativePlanner, lib  # noqa: E402
from mopa_rl_b200.model import load_model  # noqa: E402

env = sys.argv[1] if len(sys.argv) > 1 else "SawyerPushObstacle-v0"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
model = load_model(env)
ignored, passive, ref = planner_setup(model)
pl = NativePlanner(model, passive, ignored, -0.002, 0.1, seed=1)
q = random_qpos(model, n, 1234, ref)
v = pl.is_valid_host(q, flags=int(sys.argv[3]) if len(sys.argv) > 3 else 0)
out = (C.c_uint64 * 64)()
assert lib().mopa_debug_vk_stats(out) == 0
s = np.array(list(out), dtype=np.float64).reshape(16, 4)
names = ["plane_sphere", "plane_capsule", "plane_cylinder", "plane_box", "sphere_sphere", "sphere_capsule", "sphere_cylinder", "sphere_box",
         "capsule_capsule", "plane_mesh", "box_box", "mpr", "none"]
print("%s: %d queries, valid fraction %.3f, %d pairs" % (env, n, v.mean(), pl.n_pairs))
print("%-16s %10s %10s %10s %10s   (per query)" % ("class", "tested", "survive", "offending", "mpr-run"))
for i, nm in enumerate(names):
    if s[i, 0]:
        print("%-16s %10.2f %10.2f %10.3f %10.3f" % (nm, s[i, 0] / n, s[i, 1] / n, s[i, 2] / n, s[i, 3] / n))
print("%-16s %10.2f %10.2f %10.3f" % ("total", s[:, 0].sum() / n, s[:, 1].sum() / n, s[:, 2].sum() / n))
print("portal refinement: items begun %.3f per query; trips per item by phase (v1, v2, v3, refine): %s; active lanes per warp trip %.1f" % (
    s[14, 0] / n, " ".join("%.2f" % (s[13, k] / max(s[14, 0], 1)) for k in range(4)), s[14, 1] / max(s[14, 2], 1)))
