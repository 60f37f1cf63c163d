"""env.step kernel timing: tools/time_env.py [n ...] [--contacts-only]"""
import sys, time; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from mopa_rl_b200.envs import VecSawyerPushObstacle
args = [a for a in sys.argv[1:] if not a.startswith('--')]
modes = (True,) if '--contacts-only' in sys.argv else (False, True)
for n in [int(a) for a in args] or [256, 4096]:
    for contacts in modes:
        venv = VecSawyerPushObstacle(n, seed=1234, contacts=contacts)
        venv.reset()
        a = (torch.rand(n, 8, device='cuda')*2-1)
        for _ in range(2): venv.step(a)
        torch.cuda.synchronize(); t0=time.time()
        for _ in range(5): venv.step(a)
        torch.cuda.synchronize(); dt=(time.time()-t0)/5
        print('n', n, 'contacts', contacts, 'ms/step', round(dt*1e3,2), 'env-steps/s', round(n/dt), 'ncon mean', venv.ncon.float().mean().item())
