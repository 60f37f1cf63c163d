"""Stage breakdown of the env-step warp kernel (clock64 sums; needs MOPA_ENV_PROF=1) and timing for
several stage-barrier masks.  usage: python tools/env_prof.py [n_envs]"""
import ctypes as C, os, subprocess, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
if len(sys.argv) > 2 and sys.argv[2] == "child":
    import torch, numpy as np
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    n = int(sys.argv[1])
    venv = VecSawyerPushObstacle(n, seed=1234, contacts=True)
    venv.reset()
    a = torch.rand(n, 8, device="cuda") * 2 - 1
    for _ in range(3): venv.step(a)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5): venv.step(a)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 5
    print("mask %s prof %s: %.2f ms/step (%d env-steps/s)" % (os.environ.get("MOPA_ENV_SYNC_MASK"), os.environ.get("MOPA_ENV_PROF"), dt * 1e3, n / dt))
    if os.environ.get("MOPA_ENV_PROF") == "1":
        out = (C.c_uint64 * 32)()
        venv._L.mopa_env_debug_prof.argtypes = [C.c_void_p, C.c_void_p]
        venv._L.mopa_env_debug_prof(venv.h, out)
        v = np.array(list(out), dtype=np.float64)
        names = ["kinematics", "inertia+RNE", "CRBA+forces", "chol+qacc0", "contacts", "rows+PGS", "integrate"]
        tot = v[2:16].sum()
        for k in range(1, 8):
            print("  stage %d %-12s work %5.1f%%  wait %5.1f%%" % (k, names[k - 1], 100 * v[2 * k] / tot, 100 * v[2 * k + 1] / tot))
        print("  within 5: broadphase %.1f%%  narrowphase %.1f%%;  within 6: rows/Y/A %.1f%%  PGS %.1f%%  (rest = J^T f)" % tuple(100 * v[k] / tot for k in (23, 24, 25, 26)))
        tot += v[23:27].sum()
        print("  (percentages above are of the stage-clock total excluding the sub-marks; grand total incl. sub-marks %.3g cycles)" % tot)
        print("  mean broadphase survivors %.2f; substeps at the sweep cap: %d of %d" % (v[29] / max(v[22], 1), v[28], v[22]))
        sub = max(v[22], 1)
        print("  substeps with rows: %d, mean rows %.1f, mean PGS sweeps %.1f; cycles/substep/warp %.0f" % (v[22], v[21] / sub, v[20] / sub, tot / (8 * 75 * n)))
    sys.exit(0)
n = sys.argv[1] if len(sys.argv) > 1 else "4096"
MASKS = [m.split(":") for m in os.environ.get("MOPA_PROF_MASKS", "0xFE:1:14,0xFE:0:14,0x02:0:14,0xFE:0:7,0x02:0:7,0x00:0:7").split(",")]
for mask, prof, warps in MASKS:
    env = dict(os.environ, MOPA_ENV_SYNC_MASK=mask, MOPA_ENV_PROF=prof, MOPA_ENV_WARPS=warps)
    print("warps per CTA", warps, end=": ", flush=True)
    subprocess.run([sys.executable, os.path.abspath(__file__), n, "child"], env=env)
