import sys, time; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from mopa_rl_b200.envs import VecSawyerPushObstacle
from mopa_rl_b200.rollout import VecMoPARolloutRunner, MoPAConfig
n = int(sys.argv[1]) if len(sys.argv)>1 else 256
ticks = int(sys.argv[2]) if len(sys.argv)>2 else 30
venv = VecSawyerPushObstacle(n, seed=1234)
runner = VecMoPARolloutRunner(venv, MoPAConfig(max_iter=300))
torch.cuda.synchronize(); t0=time.time()
for t in range(ticks):
    runner.tick()
    if t < 3 or t % 10 == 0:
        torch.cuda.synchronize()
        print(t, 'elapsed', round(time.time()-t0,3), 'transitions', runner.n_transitions, {k:v for k,v in runner.counters.items() if v})
torch.cuda.synchronize(); dt=time.time()-t0
print('env-steps/s', runner.env_steps/dt, 'ticks/s', ticks/dt)
print('nan?', torch.isnan(venv.qpos).any().item(), 'cube z range', venv.qpos[:,29].min().item(), venv.qpos[:,29].max().item())
tr = runner.transitions[:min(runner.n_transitions, 5)].cpu().numpy()
print(tr[:, 40:52].round(3))
