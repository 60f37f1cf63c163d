import sys, time; sys.path.insert(0,'/root/repo')
import torch, numpy as np
from mopa_rl_b200.envs import VecSawyerPushObstacle
from mopa_rl_b200.rollout import VecMoPARolloutRunner, MoPAConfig
import mopa_rl_b200.rollout as R
n=4096
venv = VecSawyerPushObstacle(n, seed=1234)
runner = VecMoPARolloutRunner(venv, MoPAConfig(max_iter=1000))
for _ in range(5): runner.tick()
torch.cuda.synchronize()
# monkeypatch timers
T = {}
def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize(); t=time.perf_counter(); r=fn(*a, **k); torch.cuda.synchronize(); T[name]=T.get(name,0)+time.perf_counter()-t; return r
    return w
runner._plan = timed('plan', runner._plan)
runner._rrt_finalize = timed('rrt finalize', runner._rrt_finalize); runner._rrt_launch = timed('rrt launch', runner._rrt_launch)
runner._valid = timed('valid(in plan)', runner._valid)
venv.step = timed('env.step', venv.step)
venv.reset = timed('reset', venv.reset)
runner.policy = timed('policy', runner.policy)
t0=time.perf_counter()
K=10
for _ in range(K): runner.tick()
torch.cuda.synchronize(); tot=time.perf_counter()-t0
print('tick ms', tot/K*1e3)
for k,v in sorted(T.items(), key=lambda kv:-kv[1]): print('%-16s %.2f ms/tick'%(k, v/K*1e3))
print('other', (tot-T['plan']-T['env.step']-T.get('reset',0)-T['policy'])/K*1e3)
print(runner.counters)
