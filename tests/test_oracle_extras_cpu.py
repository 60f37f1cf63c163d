"""CPU checks of the oracle pieces added with the Newton solver, the assembly task, inverse kinematics and reuse_data:
golden fixtures (tools/make_golden.py) and analytic known answers."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def push_dyn(push_model, oracle_built):
    from mopa_rl_b200.dynmodel import DynModel

    return DynModel(push_model)


def test_assembly_env_golden(oracle_built):
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import assembly_reset_state
    from mopa_rl_b200.model import load_model
    from oracle.env_oracle import AssemblyEnvOracle

    g = np.load(os.path.join(GOLD, "assembly_env_steps.npz"))
    m = load_model("SawyerAssemblyObstacle-v0")
    dm = DynModel(m)
    n = g["qpos"].shape[1]
    q0, v0 = assembly_reset_state(m, int(g["seed"]), np.arange(n), np.zeros(n, dtype=np.int64))
    for e in range(n):
        env = AssemblyEnvOracle(m, dm)
        ob0 = env.reset_to(q0[e], v0[e])
        assert ob0.shape == (38,)
        for s in range(g["qpos"].shape[0]):
            ob, r, _ = env.step(g["actions"][s, e].astype(np.float64))
            assert np.abs(env.qpos - g["qpos"][s, e]).max() < 1e-9 and np.abs(env.qvel - g["qvel"][s, e]).max() < 1e-9
            assert np.abs(ob - g["obs"][s, e]).max() < 1e-9 and abs(r - g["reward"][s, e]) < 1e-12
        # the furniture (free body, qpos 27:34) rests on the table: it must not sink or fly
        assert abs(env.qpos[29] - q0[e][29]) < 5e-3


def test_cube_at_rest_contact_force_is_its_weight(push_model, push_dyn):
    """Known answer for the Newton solver and the get_contact_force metric: a cube at rest on the bin floor is held by
    normal forces that add up to m g; the tangential components vanish."""
    from helpers import PUSH_INIT_QPOS
    from oracle.env_oracle import PushEnvOracle

    env = PushEnvOracle(push_model, push_dyn)
    q = push_model.qpos0.copy()
    q[:7] = PUSH_INIT_QPOS
    env.reset_to(q, np.zeros(push_model.nv))
    for _ in range(3):
        env.step(np.zeros(8))
    weight = push_model.body_mass[push_model.body_name2id("cube")] * 9.81
    assert env.ncon >= 3
    assert abs(env.contact_force - weight) < 0.02 * weight, (env.contact_force, weight)


def test_ik_golden_and_jacobian(push_model, push_dyn):
    from mopa_rl_b200.envs import make_push_task
    from mopa_rl_b200.inverse_kinematics import site_frame
    from oracle.ik_oracle import IKOracle

    g = np.load(os.path.join(GOLD, "push_ik.npz"))
    task = make_push_task(push_model, push_dyn)
    body, local = site_frame(push_model, push_dyn, "grip_site")
    dofs = [int(task.arm_dof[k]) for k in range(7)]
    ik = IKOracle(push_dyn, body, local, dofs)
    for i in range(len(g["q0"])):
        for tag, tq in (("p", None), ("q", g["target_quat"][i])):
            q, err, steps, ok = ik.solve(g["q0"][i], g["target_pos"][i], tq, tol=1e-2)
            assert steps == g["steps_" + tag][i] and ok == bool(g["ok_" + tag][i])
            assert np.abs(q - g["qpos_" + tag][i]).max() < 1e-9 and abs(err - g["err_" + tag][i]) < 1e-9
            if ok:   # a converged solution puts the site within tol of the target
                assert np.linalg.norm(ik.site_pose(q)[0] - g["target_pos"][i]) < 1e-2
    # site Jacobian (a x (p - anchor)) against central differences of the forward kinematics
    q = g["q0"][0].copy()
    sp, _, joints = ik.site_pose(q)
    for c, d in enumerate(dofs):
        ax, anchor, jt = joints[d]
        col = np.cross(ax, sp - anchor)
        qa = int(push_dyn._arr["d_qadr"][d])
        qp, qm = q.copy(), q.copy()
        qp[qa] += 1e-6
        qm[qa] -= 1e-6
        num = (ik.site_pose(qp)[0] - ik.site_pose(qm)[0]) / 2e-6
        assert np.abs(col - num).max() < 1e-7, (c, col, num)


def test_reuse_data_rollout_golden(push_model, push_dyn):
    from mopa_rl_b200 import rng
    from mopa_rl_b200.rollout import MoPAConfig, planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    g = np.load(os.path.join(GOLD, "push_rollout_reuse.npz"))["records"]

    def policy(gid, k):
        u = rng.uniform01(3, np.uint64(gid), np.uint64(k), np.arange(7, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ignored, passive, _ = planner_inputs(push_model)
    cfg = MoPAConfig(max_iter=150, seed=17, reuse_data=True)
    r = ScalarMoPARunner(push_model, push_dyn, cfg, ignored, passive, 7, 2024, policy, max_episode_steps=30)
    recs = []
    for _ in range(10):
        recs.append(r.macro_step())
        recs.extend(r.extra_records)
    recs = np.array(recs, np.float32)
    assert recs.shape == g.shape and np.abs(recs - g).max() < 1e-6
    # relabelled records: planner-sized actions inside [-1, 1], intra_steps consistent with a sub-segment of a plan
    extra = recs[np.abs(recs[:, 40:47]).max(axis=1) > cfg.omega]
    assert len(extra) > 0 and np.all(np.abs(recs[:, 40:47]) <= 1.0) and np.all(recs[:, 50] >= 0)
    assert r.counters["reused"] == len(recs) - 10


def test_lift_env_golden_and_known_answers(oracle_built):
    """SawyerLiftObstacle-v0 oracle: (i) the can at rest on the bin floor is carried by a contact force equal to its
    weight, reward = reach term only; (ii) golden open -> insert can -> close -> lift sequence: has_grasp from the
    contact list, reward 0.5 = grasp_mult + (lift_mult - grasp_mult) above the lift height, + 150 inside the success
    band (env/sawyer/sawyer_lift_obstacle.py:92-148)."""
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import lift_reset_state
    from mopa_rl_b200.mjcf import mat_to_quat
    from mopa_rl_b200.model import load_model
    from oracle.env_oracle import LiftEnvOracle, _q2m

    m = load_model("SawyerLiftObstacle-v0")
    dm = DynModel(m)
    cube = dm.geoms.index(m.geom_name2id("cube"))
    assert dm._arr["g_type"][cube] == 5 and abs(dm._arr["g_size"][cube][0] - 0.0251) < 1e-4 and abs(dm._arr["g_size"][cube][1] - 0.04) < 1e-6
    g = np.load(os.path.join(GOLD, "lift_env_steps.npz"))
    n = g["qpos"].shape[1]
    q0, v0 = lift_reset_state(m, int(g["seed"]), np.arange(n), np.zeros(n, dtype=np.int64))
    q0[1, m.get_joint_qpos_addr("right_j1")] = -0.6
    ca, cva = m.get_joint_qpos_addr("cube")[0], m.get_joint_qvel_addr("cube")[0]
    weight = m.body_mass[m.body_name2id("cube")] * 9.81
    for e in range(n):
        env = LiftEnvOracle(m, dm, max_episode_steps=50)
        ob0 = env.reset_to(q0[e], v0[e])
        assert ob0.shape == (35,)
        for s in range(g["qpos"].shape[0]):
            if s == 2:   # the can appears between the open fingers, axis perpendicular to the closing direction
                tips = []
                for name in ("l_fingertip_g0", "r_fingertip_g0"):
                    gi = m.geom_name2id(name)
                    sb = dm.bodies.index(int(m.geom_bodyid[gi]))
                    tips.append(env.xpos[sb] + _q2m(env.xquat[sb]) @ m.geom_pos[gi])
                mid, cdir = 0.5 * (tips[0] + tips[1]), (tips[0] - tips[1]) / np.linalg.norm(tips[0] - tips[1])
                R = _q2m(env.xquat[env.b_ee])
                ax = R[:, 1] - (R[:, 1] @ cdir) * cdir
                ax /= np.linalg.norm(ax)
                q, v = env.qpos.copy(), env.qvel.copy()
                q[ca:ca + 3], q[ca + 3:ca + 7], v[cva:cva + 6] = mid, mat_to_quat(np.stack([cdir, np.cross(ax, cdir), ax], 1)), 0.0
                env.set_state(q, v)
                env.prev_state = None
            if env.terminal:
                break
            ob, r, d = env.step(g["actions"][s, e].astype(np.float64))
            assert np.abs(env.qpos - g["qpos"][s, e]).max() < 1e-9 and np.abs(env.qvel - g["qvel"][s, e]).max() < 1e-9
            assert np.abs(ob - g["obs"][s, e]).max() < 1e-9 and abs(r - g["reward"][s, e]) < 1e-12
            assert env.has_grasp == bool(g["grasp"][s, e])
            if s < 2:    # can at rest on the bin floor, gripper far away
                assert abs(env.contact_force - weight) < 0.02 * weight and len(env.contacts) == 1 and m.geom_name2id("cube") in env.contacts[0]
                assert r < 1e-3 and not env.has_grasp
            if env.has_grasp:
                z = ob[27]
                zt = m.body_pos[m.body_name2id("bin1")][2] + 0.45
                expect = 0.35 + (1 - np.tanh(15 * max(zt - z, 0.0))) * 0.15 + (150.0 if abs(z - zt) < 0.05 else 0.0)
                assert abs(r - expect) < 1e-9 and d == (abs(z - zt) < 0.05)
    assert g["grasp"].sum() >= 4 and (g["reward"] > 100).sum() == 1


def test_scalar_loop_discrete_and_lift_goldens(push_model, push_dyn, oracle_built):
    """Scalar restatement of MoPARolloutRunner.run with (i) config.discrete_action (omega = 0; ac_type in record slot 47,
    direct actions unscaled, relabelled records inherit ac_type) and (ii) the lift task (8-D actions, gripper entry in
    slot 47): records of one environment against the golden file."""
    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerLiftObstacle
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import MoPAConfig, env_planner_inputs, planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    g = np.load(os.path.join(GOLD, "rollout_discrete_lift.npz"))

    def pol_d(gid, k):
        u = rng.uniform01(13, np.uint64(gid), np.uint64(k), np.arange(8, dtype=np.uint64))
        return (2.0 * u[:7] - 1.0).astype(np.float32), bool(u[7] < 0.5)

    ign, pas, _ = planner_inputs(push_model)
    cfg = MoPAConfig(max_iter=150, seed=23, omega=0.0, discrete_action=True, reuse_data=True, max_reuse_data=15)
    run = ScalarMoPARunner(push_model, push_dyn, cfg, ign, pas, 300, 606, pol_d, max_episode_steps=30)
    recs = []
    for _ in range(12):
        recs.append(run.macro_step())
        recs.extend(run.extra_records)
    recs = np.array(recs, np.float32)
    assert recs.shape == g["discrete"].shape and np.abs(recs - g["discrete"]).max() < 1e-6
    assert set(np.unique(recs[:, 47])) == {0.0, 1.0}
    direct = recs[(recs[:, 47] == 0)]
    assert np.all(direct[:, 50] == 0)                                          # direct actions are single env.steps
    assert run.counters["rl"] == len(direct) and run.counters["reused"] > 0

    def pol_l(gid, k):
        u = rng.uniform01(19, np.uint64(gid), np.uint64(k), np.arange(8, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ml = load_model("SawyerLiftObstacle-v0")
    ign_l, pas_l, _ = env_planner_inputs(VecSawyerLiftObstacle, ml)
    cfg_l = MoPAConfig(max_iter=150, seed=31, reuse_data=True, max_reuse_data=15)
    run_l = ScalarMoPARunner(ml, DynModel(ml), cfg_l, ign_l, pas_l, 60, 515, pol_l, max_episode_steps=20, task="lift")
    recs_l = []
    for _ in range(8):
        recs_l.append(run_l.macro_step())
        recs_l.extend(run_l.extra_records)
    recs_l = np.array(recs_l, np.float32)
    assert recs_l.shape == g["lift"].shape and np.abs(recs_l - g["lift"]).max() < 1e-6
    assert np.all(recs_l[:, 35:40] == 0) and np.all(recs_l[:, 87:92] == 0)      # 35-float observations in 40-float rows
    main = recs_l[recs_l[:, 47] != 0]
    assert len(main) == 8                                                      # main records carry the policy's gripper entry, relabelled ones 0


def test_scalar_loop_on_the_pusher_golden(oracle_built):
    """BASELINE configs[0] on the CPU restatement: PusherObstacle-v0 MoPA-SAC loop (scripts/2d/mopa.sh: omega 0.5,
    action_range 1.0, reuse_data, max_reuse_data 30; config/pusher.py: range 0.2, contact_threshold -0.0015, step_size
    0.04) over the RK4 / PID env oracle and the SO(2) planner oracle; 4-float actions, 20-float observations."""
    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import MoPAConfig
    from oracle.rollout_oracle import ScalarMoPARunner

    g = np.load(os.path.join(GOLD, "pusher_rollout.npz"))["records"]
    m = load_model("PusherObstacle-v0")
    static = [m.geom_name2id("obstacle%d_geom" % i) for i in range(1, 8)]
    box = m.geom_name2id("box")
    ignored = [(min(box, s), max(box, s)) for s in static]
    ref = [m.get_joint_qpos_addr("joint%d" % i) for i in range(4)]
    passive = [i for i in range(m.nq) if i not in ref]
    cfg = MoPAConfig(omega=0.5, action_range=1.0, ac_scale=0.1, step_size=0.04, joint_margin=0.0, contact_threshold=-0.0015, range=0.2,
                     max_iter=1000, reuse_data=True, max_reuse_data=30, seed=5)

    def policy(gid, k):
        u = rng.uniform01(3, np.uint64(gid), np.uint64(k), np.arange(4, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    run = ScalarMoPARunner(m, DynModel(m), cfg, ignored, passive, 0, 11, policy, max_episode_steps=400, task="pusher")
    recs = []
    for _ in range(12):
        recs.append(run.macro_step())
        recs.extend(run.extra_records)
    recs = np.array(recs, np.float32)
    assert recs.shape == g.shape and np.abs(recs - g).max() < 1e-6
    assert np.all(recs[:, 44:48] == 0) and np.all(recs[:, 20:40] == 0)         # 4-float actions, 20-float observations
    assert np.allclose(recs[:, 0:4] ** 2 + recs[:, 4:8] ** 2, 1.0, atol=1e-6)   # cos / sin of the joint angles
    assert run.counters["interpolation"] + run.counters["mp"] > 0 and run.counters["reused"] > 0
    # the unlimited joint is wrapped for the planner only: joint_convert is the identity inside (-3.14, 3.14) and maps beyond
    assert abs(run._wrap(np.array([3.5, 0, 0, 0] + [0.0] * 12))[0] - (3.5 - 3.14 - 3.14)) < 1e-12
    assert run._wrap(np.array([1.0, 5.0, 0, 0] + [0.0] * 12))[1] == 5.0         # limited joints untouched
