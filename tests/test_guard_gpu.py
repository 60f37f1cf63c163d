"""Instability guard of the env-step kernel (BaseEnv._do_simulation / _after_step, env/base.py:300-304, 388-400) and the
stale contact list a lift planner-failure step reads (rl/mopa_rollouts.py:312)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_diverged_env_is_discarded_terminated_and_reset(push_model):
    import torch

    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner

    n = 32
    venv = VecSawyerPushObstacle(n, seed=3, unstable_penalty=2.5)
    venv.reset()
    q_before = venv.qpos.clone()
    # env 5: a velocity just inside MuJoCo's mjMAXVAL that blows the state past it within the step; env 9: NaN
    venv.qvel[5, 2] = 9.0e9
    venv.qvel[9, 0] = float("nan")
    act = torch.zeros(n, 8, device="cuda")
    venv.step(act)
    torch.cuda.synchronize()
    un, done, rew = venv.unstable.cpu().numpy(), venv.done.cpu().numpy(), venv.reward.cpu().numpy()
    assert un[5] == 1 and un[9] == 1 and un.sum() == 2
    assert done[5] == 1 and done[9] == 1 and done.sum() == 2
    assert rew[5] == -2.5 and rew[9] == -2.5                                        # -unstable_penalty
    assert torch.equal(venv.qpos[5], q_before[5]) and torch.equal(venv.qpos[9], q_before[9])   # the step is discarded
    assert torch.isfinite(venv.qpos).all() and torch.isfinite(venv.obs).all()
    ok = np.ones(n, bool)
    ok[[5, 9]] = False
    assert torch.isfinite(venv.qvel[torch.as_tensor(ok, device="cuda")]).all()
    assert not torch.equal(venv.qpos[0], q_before[0])                                  # healthy environments stepped

    # inside the rollout: the episode ends, the env is reset, nothing non-finite reaches the transition ring
    venv2 = VecSawyerPushObstacle(n, seed=4)
    runner = NativeMoPARolloutRunner(venv2, MoPAConfig(max_iter=50), policy=CounterPolicy(torch, venv2.dev, 1))
    for t in range(6):
        if t == 2:
            venv2.qvel[7, 1] = float("inf")
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    c = runner.counters
    assert c["unstable"] == 1 and c["episodes"] >= 1
    assert torch.isfinite(venv2.qpos).all() and torch.isfinite(venv2.qvel).all()
    rec = runner.transitions[:c["transitions"]]
    assert torch.isfinite(rec).all()
    mine = rec[rec[:, 51] == 7].cpu().numpy()
    assert mine[:, 49].sum() >= 1                                                       # a `done` record for the diverged episode
    assert int(venv2.ep_len[7]) < 6                                                     # ... and a fresh episode afterwards


def test_lift_failure_step_sees_the_previous_contact_list(oracle_built):
    """Mode-2 steps (planner failure: compute_reward without mj_step) keep the grasp of the last simulated step."""
    import torch

    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerLiftObstacle, lift_reset_state
    from mopa_rl_b200.model import load_model
    from oracle.env_oracle import LiftEnvOracle
    from test_env_gpu import _lift_fingertip_frame

    model = load_model("SawyerLiftObstacle-v0")
    n = 4
    venv = VecSawyerLiftObstacle(n, seed=13, max_episode_steps=50)
    venv.reset()
    dm = DynModel(model)
    q0, v0 = lift_reset_state(model, 13, np.arange(n), np.zeros(n, dtype=np.int64))
    envs = [LiftEnvOracle(model, dm, max_episode_steps=50) for _ in range(n)]
    for i, e in enumerate(envs):
        e.reset_to(q0[i], v0[i])
    a, va = model.get_joint_qpos_addr("cube")[0], model.get_joint_qvel_addr("cube")[0]

    def step(act, mode):
        venv.step(torch.as_tensor(act, device="cuda"), torch.as_tensor(mode, device="cuda"))
        torch.cuda.synchronize()
        out = []
        for i, e in enumerate(envs):
            if mode[i] == 2:
                r, d = e.null_step()
            else:
                _, r, d = e.step(act[i].astype(np.float64), bool(mode[i]))
            out.append(r)
        return venv.reward.cpu().numpy(), np.array(out)

    for s in range(2):                                     # open
        act = np.zeros((n, 8), np.float32)
        act[:, 7] = -1.0
        step(act, np.zeros(n, np.uint8))
    q, v = np.stack([e.qpos for e in envs]), np.stack([e.qvel for e in envs])
    for i, e in enumerate(envs):                           # the can between the fingers
        mid, quat = _lift_fingertip_frame(model, dm, e)
        q[i, a:a + 3], q[i, a + 3:a + 7], v[i, va:va + 6] = mid, quat, 0.0
        e.set_state(q[i], v[i])
        e.prev_state = None
    venv.set_state(np.arange(n), q, v)
    venv.reset_prev_state()
    for s in range(3):                                     # close
        act = np.zeros((n, 8), np.float32)
        act[:, 7] = 0.004
        g, o = step(act, np.zeros(n, np.uint8))
    assert (g >= 0.35 - 1e-9).all(), g                     # grasped
    assert np.all(venv.grasp.cpu().numpy() == 3)
    g, o = step(np.zeros((n, 8), np.float32), np.full(n, 2, np.uint8))   # planner failure: no simulation, stale contact list
    assert (g >= 0.35 - 1e-9).all() and np.abs(g - o).max() < 1e-9, (g, o)
