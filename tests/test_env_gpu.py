"""GPU <-> oracle parity of the vectorised env.step (physics + env logic), through the C ABI.
Tolerance stated by BASELINE.json north_star: qpos / qvel within 1e-5 absolute."""
from collections import OrderedDict

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _run(push_model, contacts, n=48, steps=4, planner_steps=True):
    import torch

    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerPushObstacle, push_reset_state
    from oracle.env_oracle import PushEnvOracle

    venv = VecSawyerPushObstacle(n, seed=77, contacts=contacts, max_episode_steps=3)
    venv.reset()
    dm = DynModel(push_model)
    q0, v0 = push_reset_state(push_model, 77, np.arange(n), np.zeros(n, dtype=np.int64))
    assert np.allclose(venv.qpos.cpu().numpy(), q0, atol=0, rtol=0)
    envs = [PushEnvOracle(push_model, dm, contacts=contacts, max_episode_steps=3) for _ in range(n)]
    obs0 = np.stack([e.reset_to(q0[i], v0[i]) for i, e in enumerate(envs)])
    assert np.abs(venv.obs.cpu().numpy() - obs0).max() < 1e-5
    rng = np.random.default_rng(5)
    worst = dict(qpos=0.0, qvel=0.0, obs=0.0, rew=0.0)
    for s in range(steps):
        act = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        isp = np.zeros(n, np.uint8)
        if planner_steps and s >= 1:
            isp[::2] = 1
            act[::2] *= 0.08  # planner-mode actions are joint displacements, clipped to +-ac_scale
        venv.step(torch.as_tensor(act, device="cuda"), torch.as_tensor(isp, device="cuda"))
        torch.cuda.synchronize()
        gq, gv = venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy()
        gobs, grew, gdone = venv.obs.cpu().numpy(), venv.reward.cpu().numpy(), venv.done.cpu().numpy()
        gcf = venv.cforce.cpu().numpy()
        for i, e in enumerate(envs):
            ob, r, d = e.step(act[i].astype(np.float64), bool(isp[i]))
            if contacts:   # get_contact_force metric (env/base.py:568-581)
                assert abs(gcf[i] - e.contact_force) <= 1e-6 * max(1.0, e.contact_force), (i, gcf[i], e.contact_force)
            worst["qpos"] = max(worst["qpos"], np.abs(gq[i] - e.qpos).max())
            worst["qvel"] = max(worst["qvel"], np.abs(gv[i] - e.qvel).max())
            worst["obs"] = max(worst["obs"], np.abs(gobs[i] - ob).max())
            worst["rew"] = max(worst["rew"], abs(grew[i] - r))
            assert bool(gdone[i]) == d
            assert venv.ep_len[i].item() == e.ep_len
    assert worst["qpos"] < TOL and worst["qvel"] < TOL, worst
    assert worst["obs"] < 1e-4 and worst["rew"] < 1e-6, worst   # obs is fp32 on the device side
    return worst


def test_env_step_matches_oracle_no_contacts(push_model, oracle_built):
    w = _run(push_model, contacts=False)
    print("max abs error (no contacts):", w)


def test_env_step_matches_oracle_with_contacts(push_model, oracle_built):
    w = _run(push_model, contacts=True, steps=4)
    print("max abs error (contacts):", w)


def test_assembly_env_step_matches_oracle(oracle_built):
    """SawyerAssemblyObstacle-v0 (BASELINE configs[3]): 19 simulated bodies, 39 contact geoms, peg / hole reward and the
    38-float observation, against the assembly env oracle."""
    import torch

    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerAssemblyObstacle, assembly_reset_state
    from mopa_rl_b200.model import load_model
    from oracle.env_oracle import AssemblyEnvOracle

    model = load_model("SawyerAssemblyObstacle-v0")
    n, steps = 24, 3
    venv = VecSawyerAssemblyObstacle(n, seed=31, max_episode_steps=3)
    venv.reset()
    dm = DynModel(model)
    q0, v0 = assembly_reset_state(model, 31, np.arange(n), np.zeros(n, dtype=np.int64))
    envs = [AssemblyEnvOracle(model, dm, max_episode_steps=3) for _ in range(n)]
    obs0 = np.stack([e.reset_to(q0[i], v0[i]) for i, e in enumerate(envs)])
    assert np.abs(venv.obs.cpu().numpy()[:, :38] - obs0).max() < 1e-5
    rng = np.random.default_rng(9)
    worst = dict(qpos=0.0, qvel=0.0, obs=0.0, rew=0.0)
    for s in range(steps):
        act = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        isp = np.zeros(n, np.uint8)
        if s >= 1:
            isp[::2] = 1
            act[::2] *= 0.08
        venv.step(torch.as_tensor(act, device="cuda"), torch.as_tensor(isp, device="cuda"))
        torch.cuda.synchronize()
        gq, gv = venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy()
        gobs, grew, gdone = venv.obs.cpu().numpy(), venv.reward.cpu().numpy(), venv.done.cpu().numpy()
        for i, e in enumerate(envs):
            ob, r, d = e.step(act[i].astype(np.float64), bool(isp[i]))
            worst["qpos"] = max(worst["qpos"], np.abs(gq[i] - e.qpos).max())
            worst["qvel"] = max(worst["qvel"], np.abs(gv[i] - e.qvel).max())
            worst["obs"] = max(worst["obs"], np.abs(gobs[i, :38] - ob).max())
            worst["rew"] = max(worst["rew"], abs(grew[i] - r))
            assert bool(gdone[i]) == d
    assert worst["qpos"] < TOL and worst["qvel"] < TOL, worst
    assert worst["obs"] < 1e-4 and worst["rew"] < 1e-6, worst
    print("assembly max abs error:", worst)


def _lift_fingertip_frame(model, dm, e):
    """Midpoint of the fingertip geoms, closing direction, and a can orientation whose axis is perpendicular to it."""
    from mopa_rl_b200.mjcf import mat_to_quat
    from oracle.env_oracle import _q2m

    tips = []
    for name in ("l_fingertip_g0", "r_fingertip_g0"):
        g = model.geom_name2id(name)
        sb = dm.bodies.index(int(model.geom_bodyid[g]))
        tips.append(e.xpos[sb] + _q2m(e.xquat[sb]) @ model.geom_pos[g])
    mid, cdir = 0.5 * (tips[0] + tips[1]), (tips[0] - tips[1]) / np.linalg.norm(tips[0] - tips[1])
    R = _q2m(e.xquat[e.b_ee])
    ax = R[:, 1] - (R[:, 1] @ cdir) * cdir
    ax /= np.linalg.norm(ax)
    return mid, mat_to_quat(np.stack([cdir, np.cross(ax, cdir), ax], 1))


def test_lift_env_step_matches_oracle(oracle_built):
    """SawyerLiftObstacle-v0 (BASELINE configs[2]): 8-D action with the gripper entry, 35-float observation, reward =
    max(reach, grasp, lift) with has_grasp from the contact list, success at the lift height.  The gripper is opened,
    the can is placed between the fingers, the gripper closes on it, then the arm moves: grasp / lift / success rewards are all exercised."""
    import torch

    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerLiftObstacle, lift_reset_state
    from mopa_rl_b200.model import load_model
    from oracle.env_oracle import LiftEnvOracle

    model = load_model("SawyerLiftObstacle-v0")
    n = 16
    venv = VecSawyerLiftObstacle(n, seed=13, max_episode_steps=7)
    venv.reset()
    dm = DynModel(model)
    q0, v0 = lift_reset_state(model, 13, np.arange(n), np.zeros(n, dtype=np.int64))
    assert np.array_equal(venv.qpos.cpu().numpy(), q0)
    j1 = model.get_joint_qpos_addr("right_j1")
    q0[1::4, j1] = -0.6                                   # arm lowered: the fingertips sit at the success height (bin z + 0.45)
    venv.set_state(np.arange(n), q0, v0)
    envs = [LiftEnvOracle(model, dm, max_episode_steps=7) for _ in range(n)]
    obs0 = np.stack([e.reset_to(q0[i], v0[i]) for i, e in enumerate(envs)])
    assert np.abs(venv.obs.cpu().numpy()[:, :35] - obs0).max() < 1e-5 and np.all(venv.obs.cpu().numpy()[:, 35:] == 0)
    rng = np.random.default_rng(4)
    worst = dict(qpos=0.0, qvel=0.0, obs=0.0, rew=0.0)
    kinds = set()
    a = model.get_joint_qpos_addr("cube")[0]
    va = model.get_joint_qvel_addr("cube")[0]

    def step(act, isp):
        venv.step(torch.as_tensor(act, device="cuda"), torch.as_tensor(isp, device="cuda"))
        torch.cuda.synchronize()
        gq, gv = venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy()
        gobs, grew, gdone, gsucc = venv.obs.cpu().numpy(), venv.reward.cpu().numpy(), venv.done.cpu().numpy(), venv.success.cpu().numpy()
        for i, e in enumerate(envs):
            if e.terminal:
                continue                                   # finished episodes are not stepped by the reference loop
            ob, r, d = e.step(act[i].astype(np.float64), bool(isp[i]))
            worst["qpos"] = max(worst["qpos"], np.abs(gq[i] - e.qpos).max())
            worst["qvel"] = max(worst["qvel"], np.abs(gv[i] - e.qvel).max())
            worst["obs"] = max(worst["obs"], np.abs(gobs[i, :35] - ob).max())
            worst["rew"] = max(worst["rew"], abs(grew[i] - r))
            assert bool(gdone[i]) == d and bool(gsucc[i]) == (e.success and d), (i, gdone[i], d, gsucc[i], e.success)
            kinds.add("success" if r > 100 else ("lift" if r > 0.35 + 1e-9 else ("grasp" if abs(r - 0.35) < 1e-9 else "reach")))

    for s in range(2):                                     # open the gripper (negative gripper action opens it)
        act = np.zeros((n, 8), np.float32)
        act[:, :7] = rng.uniform(-0.2, 0.2, (n, 7))
        act[:, 7] = -1.0
        step(act, np.zeros(n, np.uint8))
    q, v = np.stack([e.qpos for e in envs]), np.stack([e.qvel for e in envs])
    for i, e in enumerate(envs):                           # the can appears between the open fingers of most envs
        if i % 2 == 0 or i % 4 == 1:
            mid, quat = _lift_fingertip_frame(model, dm, e)
            q[i, a:a + 3], q[i, a + 3:a + 7], v[i, va:va + 6] = mid, quat, 0.0
        e.set_state(q[i], v[i])
        e.prev_state = None
    venv.set_state(np.arange(n), q, v)
    venv.reset_prev_state()
    for s in range(5):                                     # close gently (40 N), then move the arm with the can in hand
        act = rng.uniform(-1, 1, (n, 8)).astype(np.float32) if s >= 2 else np.zeros((n, 8), np.float32)
        act[:, 7] = 0.004
        isp = np.zeros(n, np.uint8)
        if s >= 3:
            isp[::2] = 1
            act[::2, :7] *= 0.08
        step(act, isp)
    assert worst["qpos"] < TOL and worst["qvel"] < TOL, worst
    assert worst["obs"] < 1e-4 and worst["rew"] < 1e-6, worst
    assert {"reach", "lift", "success"} <= kinds, kinds
    print("lift max abs error:", worst, sorted(kinds))


def test_lift_can_on_the_ground_plane_matches_oracle(oracle_built):
    """Plane - cylinder contacts (mjc_PlaneCylinder): the can dropped onto the ground beside the table, standing, lying on its side and
    tilted so that it lands on its rim and tumbles; kernel and oracle integrate the same trajectory."""
    import torch

    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerLiftObstacle, lift_reset_state
    from mopa_rl_b200.model import load_model
    from oracle.env_oracle import LiftEnvOracle

    model = load_model("SawyerLiftObstacle-v0")
    n = 12
    venv = VecSawyerLiftObstacle(n, seed=3, max_episode_steps=20)
    venv.reset()
    dm = DynModel(model)
    q0, v0 = lift_reset_state(model, 3, np.arange(n), np.zeros(n, dtype=np.int64))
    a = model.get_joint_qpos_addr("cube")[0]
    va = model.get_joint_qvel_addr("cube")[0]
    rng = np.random.default_rng(8)
    for i in range(n):
        q0[i, a:a + 3] = [-0.6 + 0.02 * i, 0.9, 0.06 + 0.01 * (i % 3)]
        if i % 3 == 0:
            quat = np.array([1.0, 0.0, 0.0, 0.0])
        elif i % 3 == 1:
            quat = np.array([np.sqrt(0.5), np.sqrt(0.5), 0.0, 0.0])
        else:
            quat = rng.normal(size=4)
        q0[i, a + 3:a + 7] = quat / np.linalg.norm(quat)
        v0[i, va:va + 3] = rng.uniform(-0.3, 0.3, 3)
    venv.set_state(np.arange(n), q0, v0)
    envs = [LiftEnvOracle(model, dm, max_episode_steps=20) for _ in range(n)]
    for i, e in enumerate(envs):
        e.reset_to(q0[i], v0[i])
    worst, touched = 0.0, 0
    for s in range(6):
        act = np.zeros((n, 8), np.float32)
        act[:, :7] = rng.uniform(-0.3, 0.3, (n, 7))
        venv.step(torch.as_tensor(act, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda"))
        torch.cuda.synchronize()
        gq, gv = venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy()
        for i, e in enumerate(envs):
            e.step(act[i].astype(np.float64), False)
            worst = max(worst, np.abs(gq[i] - e.qpos).max(), np.abs(gv[i] - e.qvel).max())
            touched += int(s == 5 and e.qpos[a + 2] < 0.06 and np.abs(e.qvel[va + 2]) < 0.5)
    assert touched >= 8, touched                           # the cans are held up by the ground, not falling through it
    assert min(e.qpos[a + 2] for e in envs) > 0.02
    assert worst < TOL, worst
    print("can on the ground max abs error:", worst)


def test_gym_env_view_drop_in_loop(push_model, oracle_built):
    """The reference-facing N = 1 view (gym_env.make, env/__init__.py:7-32 + BaseEnv API) driven the way
    rl/trainer.py:62-75 and rl/mopa_rollouts.py drive the reference env - trainer-style planner set-up from the env's
    attributes, a direct step, planner steps through form_action, the planner-failure pair compute_reward /
    _after_step - against the env oracle."""
    import types

    from mopa_rl_b200 import gym_env
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import push_reset_state
    from mopa_rl_b200.motion_planners import SamplingBasedPlanner
    from oracle.env_oracle import PushEnvOracle

    env = gym_env.make("SawyerPushObstacle-v0", seed=77, max_episode_steps=12)
    ob = env.reset()
    q0, v0 = push_reset_state(push_model, 77, [0], [0])
    orc = PushEnvOracle(push_model, DynModel(push_model), max_episode_steps=12)
    ob0 = orc.reset_to(q0[0], v0[0])
    cat = lambda o: np.concatenate(list(o.values()))
    assert list(ob.keys())[:3] == ["joint_pos", "joint_vel", "gripper_qpos"] and np.abs(cat(ob) - ob0).max() < 1e-5
    # rl/trainer.py:62-75: ignored contacts and passive joints from the env's attributes
    ignored = [(min(a, b), max(a, b)) for a in env.manipulation_geom_ids for b in env.static_geom_ids]
    passive = [i for i in range(len(env.sim.data.qpos)) if i not in env.ref_joint_pos_indexes]
    cfg = types.SimpleNamespace(planner_type="rrt_connect", range=0.1, planner_objective="path_length", threshold=0.0, seed=1234)
    planner = SamplingBasedPlanner(cfg, env.xml_path, env.sim.model.nu, [], passive_joint_idx=passive, ignored_contacts=ignored,
                                   contact_threshold=-0.002)
    assert planner.isValidState(env.sim.data.qpos)
    # direct step
    a = np.random.default_rng(1).uniform(-1, 1, 7).astype(np.float32)
    ob, r, d, info = env.step(OrderedDict([("default", a)]))
    o2, r2, d2 = orc.step(a.astype(np.float64))
    assert np.abs(cat(ob) - o2).max() < 1e-4 and abs(r - r2) < 1e-9 and d == d2 and info == {}
    assert np.abs(env.sim.data.qpos - orc.qpos).max() < 1e-5 and env.sim.data.ncon == orc.ncon
    assert abs(env.get_contact_force() - orc.contact_force) < 1e-6 * max(1.0, orc.contact_force)
    # a short plan executed waypoint by waypoint (rl/mopa_rollouts.py:166-199)
    curr = env.sim.data.qpos.copy()
    target = curr.copy()
    target[env.ref_joint_pos_indexes] -= 0.03                      # (+0.03 on all joints puts a forearm geom 2.8 mm into the pedestal)
    traj, _, valid, exact = planner.plan(curr, target, timelimit=0.3)
    assert valid and (not exact or len(traj) >= 2)
    for nq in (traj[1:] if exact else []):
        ac = env.form_action(nq)
        ob, r, d, info = env.step(ac, is_planner=True)
        o2, r2, d2 = orc.step(np.asarray(ac["default"], np.float32).astype(np.float64), True)
        assert np.abs(cat(ob) - o2).max() < 1e-4 and abs(r - r2) < 1e-9 and d == d2
        if d:
            break
    env._reset_prev_state()
    orc.prev_state = None
    # planner failure: compute_reward(zeros) + _after_step (rl/mopa_rollouts.py:304-327)
    if not env._terminal:
        reward, info = env.compute_reward(np.zeros(env.sim.model.nu))
        done, info, _ = env._after_step(reward, False, info)
        r2, d2 = orc.null_step()
        assert abs(reward - r2) < 1e-9 and done == d2 and env._episode_length == orc.ep_len
    grip = env.sim.data.get_site_xpos("grip_site")
    assert np.abs(grip - (orc._site(orc.b_ee, orc.s_grip))).max() < 1e-3      # frames of the last substep vs current qpos
    env.close()
