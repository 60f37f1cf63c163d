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


def test_native_runner_on_the_pusher_matches_scalar_reference_loop(oracle_built):
    """BASELINE configs[0] through the vectorised runner: PusherObstacle-v0 with the 2-D preset (scripts/2d/mopa.sh: omega 0.5,
    action_range 1.0, reuse_data, max_reuse_data 30; config/pusher.py: range 0.2 / 0.1, contact_threshold -0.0015, step_size 0.04),
    4-D actions, the unlimited joint0 wrapped for the planner and un-wrapped when its path is re-based, rejection-sampled resets."""
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecPusherObstacle
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, env_planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    model = load_model("PusherObstacle-v0")
    n, ticks, seed, off = 16, 90, 11, 224   # env 235 is one of the few whose RRT-Connect problem is solvable (2 of 400 envs in 14 macro actions)
    cfg = MoPAConfig(omega=0.5, action_range=1.0, ac_scale=0.1, step_size=0.04, joint_margin=0.0, contact_threshold=-0.0015, range=0.2,
                     simple_planner_range=0.1, max_iter=1000, simple_max_iter=6, reuse_data=True, max_reuse_data=30, seed=5, debug_block_mod=2)
    venv = VecPusherObstacle(n, seed=seed, max_episode_steps=12, env_id_offset=off)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 3, action_dim=4))
    for _ in range(ticks):
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    c = runner.counters
    rec = runner.transitions[:c["transitions"]].cpu().numpy()
    assert c["episodes"] >= n and c["interpolation"] > n and c["reused"] > 0 and c["mp"] > 0, c

    def policy(gid, k):
        u = rng.uniform01(3, np.uint64(gid), np.uint64(k), np.arange(4, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ignored, passive, _ = env_planner_inputs(VecPusherObstacle, model)
    dm = DynModel(model)
    worst, n_plan = 0.0, 0
    for e in range(n):
        gid = off + e
        mine = list(rec[rec[:, 51] == gid])
        ref = ScalarMoPARunner(model, dm, cfg, ignored, passive, gid, seed, policy, max_episode_steps=12, task="pusher")
        k = 0
        while k < len(mine):
            for o in [ref.macro_step()] + list(ref.extra_records):
                if k >= len(mine):
                    break
                r = mine[k]
                assert np.allclose(r[40:44], o[40:44], atol=1e-6) and np.all(r[44:48] == 0), (e, k, r[40:48], o[40:48])
                assert r[49] == o[49] and r[50] == o[50], (e, k, r[48:51], o[48:51])
                assert abs(r[48] - o[48]) < 1e-5, (e, k)
                d = max(np.abs(r[0:20] - o[0:20]).max(), np.abs(r[52:72] - o[52:72]).max())
                worst = max(worst, d)
                assert d < 1e-4, (e, k, d)
                n_plan += r[50] > 0
                k += 1
    assert n_plan > n
    print("pusher: native vs scalar runner: %d records, worst |obs diff| %.2e, counters %s" % (len(rec), worst, c))
