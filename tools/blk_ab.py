import os, subprocess, sys
if len(sys.argv) > 1:
    sys.path.insert(0, '/root/repo')
    import torch, numpy as np
    from mopa_rl_b200 import envs
    cls = getattr(envs, sys.argv[2])
    n = 256
    venv = cls(n, seed=5)
    venv.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(3)
    for k in range(40):
        a = torch.rand(n, 8, device='cuda', generator=g) * 2 - 1
        venv.step(a)
    torch.cuda.synchronize()
    np.save(sys.argv[1], np.concatenate([venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy()], 1))
    sys.exit(0)
for cls in ("VecSawyerPushObstacle", "VecSawyerLiftObstacle", "VecSawyerAssemblyObstacle"):
    for mask, f in (("0xFE", "/tmp/ab_blk.npy"), ("0x2FE", "/tmp/ab_dense.npy")):
        subprocess.run([sys.executable, __file__, f, cls], env=dict(os.environ, MOPA_ENV_SYNC_MASK=mask), check=True)
    import numpy as np
    a, b = np.load("/tmp/ab_blk.npy"), np.load("/tmp/ab_dense.npy")
    d = np.abs(a - b)
    print(cls, "bitwise equal:", np.array_equal(a, b), "max abs diff", d.max(), "envs differing", int((d.max(1) > 0).sum()))
