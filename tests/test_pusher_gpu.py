"""PusherObstacle-v0 on the GPU (BASELINE configs[0]; env/pusher/pusher_obstacle.py): RK4 + PID env.step, the rejection-sampled
reset and the observation / reward, against PusherEnvOracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_pusher_reset_and_step_match_oracle(oracle_built):
    import torch

    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecPusherObstacle
    from mopa_rl_b200.model import load_model
    from oracle.env_oracle import PusherEnvOracle

    model = load_model("PusherObstacle-v0")
    n, seed = 12, 404
    venv = VecPusherObstacle(n, seed=seed, env_id_offset=30, max_episode_steps=5)
    obs0 = venv.reset().cpu().numpy()
    dm = DynModel(model)
    envs = [PusherEnvOracle(model, dm, max_episode_steps=5) for _ in range(n)]
    ob_ref = np.stack([e.reset(seed, 30 + i, 0) for i, e in enumerate(envs)])
    gq, gv = venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy()
    for i, e in enumerate(envs):      # the same accepted draw (same attempt), the same observation
        assert np.array_equal(gq[i], e.qpos) and np.array_equal(gv[i], e.qvel), i
    assert np.abs(obs0[:, :20] - ob_ref).max() < 1e-6 and np.all(obs0[:, 20:] == 0)
    rng = np.random.default_rng(8)
    worst = dict(qpos=0.0, qvel=0.0, obs=0.0, rew=0.0)
    for s in range(5):
        act = rng.uniform(-1, 1, (n, 4)).astype(np.float32) * (0.1 if s else 0.6)   # a large first move, then planner-sized steps
        isp = np.zeros(n, np.uint8)
        if s >= 2:
            isp[::2] = 1
        if s == 3:
            isp[1] = 2                                     # a planner-failure step: reward / accounting without simulation
        venv.step(torch.as_tensor(act, device="cuda"), torch.as_tensor(isp, device="cuda"))
        torch.cuda.synchronize()
        gq, gv = venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy()
        gobs, grew, gdone = venv.obs.cpu().numpy(), venv.reward.cpu().numpy(), venv.done.cpu().numpy()
        for i, e in enumerate(envs):
            if isp[i] == 2:
                r, d = e.null_step()
                ob = None
            else:
                ob, r, d = e.step(act[i].astype(np.float64), bool(isp[i]))
            worst["qpos"] = max(worst["qpos"], np.abs(gq[i] - e.qpos).max())
            worst["qvel"] = max(worst["qvel"], np.abs(gv[i] - e.qvel).max())
            if ob is not None:
                worst["obs"] = max(worst["obs"], np.abs(gobs[i, :20] - ob).max())
            worst["rew"] = max(worst["rew"], abs(grew[i] - r))
            assert bool(gdone[i]) == d, (s, i)
    assert bool(gdone.all())                               # max_episode_steps = 5
    assert worst["qpos"] < TOL and worst["qvel"] < TOL, worst
    assert worst["obs"] < 1e-5 and worst["rew"] < 1e-6, worst
    # the next episode draws a new state, again the oracle's
    venv.reset(np.arange(0, n, 3))
    gq = venv.qpos.cpu().numpy()
    for i in range(0, n, 3):
        envs[i].reset(seed, 30 + i, 1)
        assert np.array_equal(gq[i], envs[i].qpos), i
    print("pusher env.step (100 RK4 mj_steps, PID): max abs error", worst)
