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
