"""Physics-step oracle: rigid-body identities, closed-form cases, contacts, env logic, goldens."""
import os

import numpy as np
import pytest

from helpers import PUSH_INIT_QPOS

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def dyn(push_model, oracle_built):
    from mopa_rl_b200.dynmodel import DynModel

    dm = DynModel(push_model)
    return dm, oracle_built.OracleDyn(dm)


def test_simulated_subtrees(dyn, push_model):
    dm, od = dyn
    names = [push_model.names["body"][b] for b in dm.bodies]
    assert names == ["right_l0", "head", "screen", "right_l1", "right_l2", "right_l3", "right_l4", "right_l5", "right_l6",
                     "right_ee_attchment", "clawGripper", "rightclaw", "leftclaw", "cube"]
    assert dm.nd == 15 and dm.nact == 7 and len(dm.pairs) == 250
    assert not any("indicator" in n or "target" in n for n in names)        # ghost arms / target slider are not integrated


def _integrate(m, dm, q, v, eps):
    from mopa_rl_b200.mjcf import quat_mul

    q2 = q.copy()
    for qa, va in zip(dm.dof_qadr, dm.dof_vadr):
        if qa >= 0:
            q2[qa] += eps * v[va]
    a, va = m.get_joint_qpos_addr("cube")[0], m.get_joint_qvel_addr("cube")[0]
    w = v[va + 3:va + 6]
    ang = np.linalg.norm(w) * abs(eps)
    if ang > 0:
        ax = w / np.linalg.norm(w) * np.sign(eps)
        q2[a + 3:a + 7] = quat_mul(q2[a + 3:a + 7], np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax]))
    return q2


def test_mass_matrix_and_bias_identities(dyn, push_model):
    m = push_model
    dm, od = dyn
    rng = np.random.default_rng(0)
    q = m.qpos0.copy()
    q[:7] = PUSH_INIT_QPOS + rng.uniform(-0.5, 0.5, 7)
    v = np.zeros(m.nv)
    v[dm.dof_vadr] = rng.uniform(-1, 1, dm.nd)
    M, bias, com = od.mass_bias(q, v)
    assert np.abs(M - M.T).max() == 0 and np.linalg.eigvalsh(M).min() > 0
    mass = m.body_mass[dm.bodies]
    eps = 1e-6
    # gravity torque = dV/dq (bias at zero velocity)
    _, b0, _ = od.mass_bias(q, np.zeros(m.nv))

    def V(qq):
        return 9.81 * np.sum(mass * od.mass_bias(qq, np.zeros(m.nv))[2][:, 2])

    g = np.zeros(dm.nd)
    for k in range(dm.nd):
        e = np.zeros(m.nv)
        e[dm.dof_vadr[k]] = 1
        g[k] = (V(_integrate(m, dm, q, e, eps)) - V(_integrate(m, dm, q, e, -eps))) / (2 * eps)
    assert np.abs(g - b0).max() < 1e-6
    # Coriolis forces: qd . C(q, qd) = 1/2 qd^T Mdot qd
    qd = v[dm.dof_vadr]
    M1 = od.mass_bias(_integrate(m, dm, q, v, eps), v)[0]
    M0 = od.mass_bias(_integrate(m, dm, q, v, -eps), v)[0]
    assert abs(qd @ (bias - b0) - 0.5 * qd @ ((M1 - M0) / (2 * eps)) @ qd) < 1e-6
    # kinetic energy from finite-difference COM velocities (translation part) is bounded by 1/2 qd^T M qd
    c1 = od.mass_bias(_integrate(m, dm, q, v, eps), v)[2]
    c0 = od.mass_bias(_integrate(m, dm, q, v, -eps), v)[2]
    T_lin = 0.5 * np.sum(mass * np.sum(((c1 - c0) / (2 * eps)) ** 2, axis=1))
    assert 0 < T_lin < 0.5 * qd @ (M - np.diag(m.dof_armature[dm.dof_vadr])) @ qd


def test_free_fall_and_hold(dyn, push_model):
    m = push_model
    dm, od = dyn
    od.enable_contacts(0)
    q = m.qpos0.copy()
    q[:7] = PUSH_INIT_QPOS
    v = np.zeros(m.nv)
    bias, _, _ = od.forward(q, v)
    comp = np.zeros(dm.nd, np.int32)
    comp[:7] = 1
    n = 50
    q1, v1, _, _, _, _ = od.step(q, v, q[:7], comp, bias, n)
    t = n * m.opt_timestep
    # cube in free fall (semi-implicit Euler: z = z0 - g h^2 n(n+1)/2), tiny free-joint damping 5e-4 ignored at 1e-4
    assert abs(q1[29] - (0.88 - 9.81 * m.opt_timestep ** 2 * n * (n + 1) / 2)) < 1e-4
    assert abs(v1[m.get_joint_qvel_addr("cube")[0] + 2] + 9.81 * t) < 1e-3
    # arm held by the position actuators + gravity compensation: stays where it is
    assert np.abs(q1[:7] - q[:7]).max() < 2e-4
    od.enable_contacts(1)


def test_actuator_tracking_and_joint_limits(dyn, push_model):
    m = push_model
    dm, od = dyn
    q = m.qpos0.copy()
    q[:7] = PUSH_INIT_QPOS
    v = np.zeros(m.nv)
    bias, _, _ = od.forward(q, v)
    comp = np.zeros(dm.nd, np.int32)
    comp[:7] = 1
    target = q[:7] + np.array([0.05, -0.05, 0.05, 0.05, -0.05, 0.05, -0.05])
    q1, v1, bias, _, _, _ = od.step(q, v, target, comp, bias, 75)
    moved = (q1[:7] - q[:7]) / (target - q[:7])
    assert (moved > 0.15).all() and (moved < 1.1).all()                    # heads to the target, no overshoot beyond 10 %
    q2, v2, bias, _, _, _ = od.step(q1, v1, target, comp, bias, 600)
    assert np.abs(q2[:7] - target).max() < 5e-3 and np.abs(v2[:7]).max() < 1e-2   # settles on the target
    # drive right_j1 (range [-3.8, 1.25]) far past its upper limit: the soft limit holds it near 1.25
    q3 = q.copy()
    q3[1] = 1.2
    tgt = q3[:7].copy()
    tgt[1] = 3.0
    b3, _, _ = od.forward(q3, v)
    q4, _, _, _, _, _ = od.step(q3, v, tgt, comp, b3, 400)
    assert 1.2 < q4[1] < 1.25 + 0.05


def test_cube_rests_on_the_bin_floor(dyn, push_model):
    m = push_model
    dm, od = dyn
    q = m.qpos0.copy()
    q[:7] = PUSH_INIT_QPOS
    v = np.zeros(m.nv)
    bias, _, _ = od.forward(q, v)
    comp = np.zeros(dm.nd, np.int32)
    comp[:7] = 1
    q1, v1, bias, _, _, ncon = od.step(q, v, q[:7], comp, bias, 300)
    cv = m.get_joint_qvel_addr("cube")[0]
    assert ncon == 4                                                        # box-on-box face contact: four clipped corners
    assert abs(q1[29] - 0.86) < 1e-3 and np.abs(v1[cv:cv + 6]).max() < 1e-3  # bin floor top 0.83 + half cube 0.03, sub-mm sink
    assert np.abs(q1[27:29] - [0.92, 0.0]).max() < 1e-3                      # does not slide
    assert abs(np.linalg.norm(q1[30:34]) - 1) < 1e-12


def test_golden_env_steps(dyn, push_model):
    from mopa_rl_b200.envs import push_reset_state
    from oracle.env_oracle import PushEnvOracle

    g = np.load(os.path.join(GOLD, "push_env_steps.npz"))
    n = g["actions"].shape[1]
    q0, v0 = push_reset_state(push_model, int(g["seed"]), np.arange(n), np.zeros(n, dtype=np.int64))
    for e in range(n):
        env = PushEnvOracle(push_model, dyn[0])
        ob = env.reset_to(q0[e], v0[e])
        assert ob.shape == (40,)
        for s in range(3):
            ob, r, d = env.step(g["actions"][s, e].astype(np.float64))
            assert np.abs(env.qpos - g["qpos"][s, e]).max() < 1e-9 and np.abs(env.qvel - g["qvel"][s, e]).max() < 1e-9
            assert abs(r - g["reward"][s, e]) < 1e-12 and not d


def test_env_logic(dyn, push_model):
    from mopa_rl_b200.envs import push_reset_state
    from oracle.env_oracle import PushEnvOracle

    q0, v0 = push_reset_state(push_model, 3, [0], [0])
    env = PushEnvOracle(push_model, dyn[0], max_episode_steps=3)
    ob = env.reset_to(q0[0], v0[0])
    # observation layout (SawyerEnv._get_obs + push _get_obs): 7+7+2+2+3+4+3+3+4+3+2
    assert np.allclose(ob[:7], q0[0][:7]) and np.allclose(ob[7:14], 0)
    assert np.allclose(ob[35:38], ob[18:21] - ob[28:31]) and np.allclose(ob[38:40], ob[28:30] - ob[25:27])
    assert np.isclose(np.linalg.norm(ob[21:25]), 1) and np.allclose(ob[31:35], [0, 0, 0, 1])      # quats xyzw
    # direct step latches prev_state from the live qpos, clips the action to +-1 * ac_scale
    env.step(np.full(7, 5.0))
    assert np.allclose(env.prev_state, q0[0][:7] + 0.05)
    # planner steps keep integrating the latched desired state and clip the displacement to +-ac_scale
    live = env.qpos[:7].copy()
    env.step(np.full(7, 0.2), is_planner=True)
    assert np.allclose(env.prev_state, q0[0][:7] + 0.10) and not np.allclose(env.prev_state, live + 0.05)
    _, _, done = env.step(np.zeros(7))
    assert done and env.ep_len == 3                                                                   # max_episode_steps
    # success: cube within distance_threshold of the target -> +150 and terminal
    q = q0[0].copy()
    q[27:29] = push_model.body_pos[push_model.body_name2id("target")][:2] + q[34:36] + [0.03, 0.0]
    q[29] = 0.86
    env2 = PushEnvOracle(push_model, dyn[0])
    env2.reset_to(q, v0[0])
    _, r, done = env2.step(np.zeros(7))
    assert done and env2.success and r > 150


# ---------------------------------------------------------------------------- Pusher (BASELINE configs[0]): RK4, velocity actuators
@pytest.fixture(scope="module")
def pusher(oracle_built):
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.model import load_model

    m = load_model("PusherObstacle-v0")
    return m, DynModel(m)


def test_pusher_dynmodel_splits_multi_joint_bodies(pusher):
    """The box and the target carry two slide joints each: one simulated body per joint, massless virtual bodies first."""
    m, dm = pusher
    assert dm.nb == 9 and dm.nd == 8 and dm.nact == 4 and m.opt_integrator == 1
    A = dm._arr
    virt = [i for i, b in enumerate(dm.bodies) if b < 0]
    assert len(virt) == 2 and all(A["b_mass"][i] == 0 and A["b_jtype"][i] == 2 for i in virt)
    for i in virt:                                   # the real body hangs off its virtual parent with an identity offset
        assert A["b_parent"][i + 1] == i and np.all(A["b_pos"][i + 1] == 0) and dm.bodies[i + 1] == -1 - dm.bodies[i]
    from oracle.oracle import OracleDyn

    M, bias, _ = OracleDyn(dm).mass_bias(m.qpos0, np.zeros(m.nv))
    box_mass = 1000.0 * 0.02 ** 3                    # inertiafromgeom="true": the <inertial> element of the box is overridden
    assert np.allclose(np.diag(M)[-2:], box_mass) and np.all(np.abs(bias) < 1e-12)   # planar arm, gravity along the hinge axes
    assert np.all(np.diag(M)[:4] > 1.0)              # armature 1 + link inertias


def test_pusher_rk4_step_is_the_rk4_polynomial_of_the_linearised_system(pusher):
    """Known answer for mj_RungeKutta (N = 4): with tiny velocities the arm is the linear system M qdd = -(d + kv gear^2) qd
    (joint damping 1, velocity actuators kv 1 x gear 10 with ctrl 0), for which one RK4 step multiplies qd by
    I + Z + Z^2/2 + Z^3/6 + Z^4/24, Z = h M^-1 (-(d + kv gear^2))."""
    from oracle.oracle import OracleDyn

    m, dm = pusher
    o = OracleDyn(dm)
    q, v = m.qpos0.copy(), np.zeros(m.nv)
    v[:4] = [1e-6, -2e-6, 3e-6, -1e-6]
    M, _, _ = o.mass_bias(q, v)
    Z = -np.linalg.solve(M[:4, :4], np.eye(4) * (1.0 + 1.0 * 10.0 ** 2)) * m.opt_timestep
    P = np.eye(4) + Z + Z @ Z / 2 + Z @ Z @ Z / 6 + Z @ Z @ Z @ Z / 24
    v1 = o.step(q, v, np.zeros(4), np.zeros(dm.nd, np.int32), np.zeros(dm.nd), 1)[1]
    assert np.abs(v1[:4] - P @ v[:4]).max() < 1e-15 * 1e3 * np.abs(v[:4]).max() + 1e-20
    assert np.all(v1[4:] == 0)


def test_pusher_env_oracle_tracks_pushes_and_matches_golden(pusher):
    import os

    from oracle.env_oracle import PusherEnvOracle

    m, dm = pusher
    env = PusherEnvOracle(m, dm)
    # reset: the acceptance test of PusherObstacleEnv._reset holds for the accepted draw
    ob = env.reset(7, 0, 0)
    assert ob.shape == (20,) and env.ncon_at(env.qpos) == 0
    goal, box = env.qpos[-4:-2], env.qpos[-2:]
    assert goal[0] <= box[0] and np.linalg.norm(env.xpos[env.b_box] - env.xpos[env.b_target]) > 0.1
    assert np.allclose(ob[:4] ** 2 + ob[4:8] ** 2, 1.0) and np.allclose(ob[18:20], goal)
    # free motion: the PID + velocity-actuator loop tracks prev_state + action within a few hundredths of a radian per 1 s step
    a = np.array([0.05, -0.08, 0.1, -0.03])
    start = env.qpos[env.ref_q].copy()
    env.step(a)
    assert env.ncon == 0 and np.abs(env.qpos[env.ref_q] - (start + a)).max() < 0.03 and np.allclose(env.qpos[-2:], box)
    # pushing: the stretched arm sweeps into the box, the box moves, reward_reach is paid (fingertip within 0.1 of the box)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pusher_env_steps.npz"))
    env.reset_to(g["qpos0"], np.zeros(m.nv))
    moved = False
    for s in range(len(g["actions"])):
        ob, r, d = env.step(g["actions"][s])
        assert np.abs(env.qpos - g["qpos"][s]).max() < 1e-9 and np.abs(env.qvel - g["qvel"][s]).max() < 1e-8
        assert np.abs(ob - g["obs"][s]).max() < 1e-9 and abs(r - g["reward"][s]) < 1e-12 and env.ncon == g["ncon"][s]
        moved = moved or np.abs(env.qpos[-2:] - g["qpos0"][-2:]).max() > 0.01
        assert 0.0 < r < 0.1 and not d
    assert moved and g["ncon"].max() >= 1


# ---------------------------------------------------------------------------- plane - cylinder contacts (the lift can on the ground plane)
def test_can_rests_on_the_ground_plane_upright_and_on_its_side(oracle_built):
    """engine_collision_primitive.c mjc_PlaneCylinder: up to three points under the rim nearest the plane when the can stands (disc contact),
    two along the lowest generator when it lies on its side.  Known answers: the can's centre settles at half height / radius above the
    ground (sub-mm sink), it keeps its orientation and it stops."""
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.model import load_model
    from oracle.oracle import OracleDyn

    m = load_model("SawyerLiftObstacle-v0")
    dm = DynModel(m)
    A = dm._arr
    gi = [i for i, g in enumerate(dm.geoms) if g == m.geom_name2id("cube")][0]
    assert int(A["g_type"][gi]) == 5
    r, hh = float(A["g_size"][gi][0]), float(A["g_size"][gi][1])
    off = float(A["g_pos"][gi][2])
    a = m.get_joint_qpos_addr("cube")[0]
    va = m.get_joint_qvel_addr("cube")[0]
    od = OracleDyn(dm)
    for quat, rest, ncon_want in (([1.0, 0.0, 0.0, 0.0], hh - off, 3), ([np.sqrt(0.5), np.sqrt(0.5), 0.0, 0.0], r, 2)):
        q = m.qpos0.copy()
        v = np.zeros(m.nv)
        q[a:a + 3] = [-0.6, 0.9, rest + 0.05]        # clear of the table and the robot pedestal
        q[a + 3:a + 7] = quat
        bias = np.zeros(dm.nd)
        comp = np.zeros(dm.nd, np.int32)
        ctrl = np.array([q[int(dm.dof_qadr[int(k)])] for k in A["a_dof"]])   # position actuators hold the arm where it is
        for _ in range(8):
            q, v, bias, _, _, ncon = od.step(q, v, ctrl, comp, bias, 75)
        assert ncon == ncon_want, (quat, ncon)
        assert rest - 1.5e-3 < q[a + 2] < rest + 1e-4, (quat, q[a + 2], rest)
        assert np.abs(v[va:va + 3]).max() < 1e-3
        assert abs(abs(np.dot(q[a + 3:a + 7], quat)) - 1.0) < 1e-3           # neither tips over nor rolls away


# ---------------------------------------------------------------------------- use_ik_target: mat2quat, the two routes
def test_mat2quat_eigen_route_equals_the_closed_form_up_to_float32_resolution():
    """util.env.mat2quat (util/env.py:232-289) takes the eigenvector of a symmetric 4x4 matrix built from the float32 copy of the
    rotation and fixes the sign with w >= 0; the device IK front end (csrc/ik.cu, rollout mode) uses mju_mat2Quat's closed form on
    the same float32 copy and the same sign rule.  Both are the unit quaternion of the rotation: they must agree to the resolution
    of the float32 matrix, also near w = 0 where the sign rule flips (the target orientation is sign invariant there)."""
    from oracle.ik_oracle import _mat2quat, _q2m

    rng = np.random.default_rng(3)
    worst = 0.0
    for k in range(400):
        q = rng.normal(size=4)
        if k % 10 == 0:
            q[0] = 1e-4 * rng.normal()                      # half-turn rotations: w close to 0
        q /= np.linalg.norm(q)
        R = _q2m(q)
        M = np.array(R, dtype=np.float32)
        m00, m01, m02, m10, m11, m12, m20, m21, m22 = [float(x) for x in M.ravel()]
        K = np.array([[m00 - m11 - m22, 0.0, 0.0, 0.0], [m01 + m10, m11 - m00 - m22, 0.0, 0.0],
                      [m02 + m20, m12 + m21, m22 - m00 - m11, 0.0], [m21 - m12, m02 - m20, m10 - m01, m00 + m11 + m22]]) / 3.0
        w, V = np.linalg.eigh(K)
        qe = V[[3, 0, 1, 2], np.argmax(w)]
        qe = -qe if qe[0] < 0 else qe                        # (w, x, y, z)
        qc = _mat2quat(M.astype(np.float64))
        qc = -qc if qc[0] < 0 else qc
        d = min(np.abs(qe - qc).max(), np.abs(qe + qc).max())   # same rotation; the sign may differ only when w ~ 0
        assert np.abs(qe - qc).max() < 2e-6 or abs(qe[0]) < 1e-3, (k, qe, qc)
        worst = max(worst, d)
    assert worst < 2e-6, worst
