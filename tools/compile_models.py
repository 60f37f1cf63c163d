"""Compile the reference's MJCF scenes into flat-array .npz files (mopa_rl_b200/assets).

The XML + STL assets live in the reference checkout (/root/reference/env/assets) and are an
INPUT FORMAT, not source we ship; the GPU box has no reference checkout, so the compiled
numeric scenes are committed as generated fixtures.  Re-run after changing mjcf.py:

    python tools/compile_models.py [/root/reference]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mopa_rl_b200.mjcf import compile_mjcf  # noqa: E402
from mopa_rl_b200.model import ASSET_DIR  # noqa: E402

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
os.makedirs(ASSET_DIR, exist_ok=True)
for name in ["sawyer_push_obstacle", "sawyer_lift_obstacle", "sawyer_assembly_obstacle", "pusher_obstacle"]:
    m = compile_mjcf(os.path.join(ref, "env", "assets", "xml", name + ".xml"))
    out = os.path.join(ASSET_DIR, name + ".npz")
    m.save(out)
    print(name, "nq", m.nq, "nv", m.nv, "nbody", m.nbody, "ngeom", m.ngeom, "->", out, os.path.getsize(out), "bytes")
