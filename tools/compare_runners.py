import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, time
from mopa_rl_b200 import rng
from mopa_rl_b200.dynmodel import DynModel
from mopa_rl_b200.envs import VecSawyerPushObstacle
from mopa_rl_b200.model import load_model
from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, VecMoPARolloutRunner, planner_inputs
n, ticks, seed = 12, 60, 4321
cfg = MoPAConfig(max_iter=150, seed=99)
class DensePolicy(CounterPolicy): pass
venv = VecSawyerPushObstacle(n, seed=seed, max_episode_steps=25, env_id_offset=100)
runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 7))
for _ in range(ticks): runner.tick()
runner.drain(); torch.cuda.synchronize()
print('native counters', runner.counters)
nt = runner.n_transitions
rec = runner.transitions[:nt].cpu().numpy()
venv2 = VecSawyerPushObstacle(n, seed=seed, max_episode_steps=25, env_id_offset=100)
r2 = VecMoPARolloutRunner(venv2, cfg, policy=CounterPolicy(torch, venv2.dev, 7))
for _ in range(ticks): r2.tick()
r2.drain(); torch.cuda.synchronize()
print('torch counters', r2.counters, r2.env_steps, r2.n_transitions)
rec2 = r2.transitions[:r2.n_transitions].cpu().numpy()
worst=0; cmp=0
for e in range(n):
    a = rec[rec[:,51]==100+e]; b = rec2[rec2[:,51]==100+e]
    k = min(len(a),len(b))
    for i in range(k):
        d = np.abs(a[i]-b[i]).max(); worst=max(worst,d); cmp+=1
        if d > 1e-4: print('MISMATCH env',e,'rec',i,d, a[i][40:52], b[i][40:52]); break
print('compared',cmp,'records, worst diff',worst)
