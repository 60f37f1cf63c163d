"""MJCF-subset compiler: compiled-model facts and MuJoCo compile rules."""
import os

import numpy as np
import pytest

from mopa_rl_b200.model import load_model

REF_XML = "/root/reference/env/assets/xml"

APPENDIX_A = {
    "sawyer_push_obstacle": dict(nq=36, nv=35, nu=7, njnt=30, nbody=50, ngeom=87, collidable=27, by_type={0: 1, 6: 11, 5: 8, 2: 3, 3: 4}),
    "sawyer_lift_obstacle": dict(nq=34, nv=33, nu=9, njnt=28, nbody=61, ngeom=100, collidable=29, by_type={0: 1, 6: 14, 5: 6, 2: 3, 3: 4, 7: 1}),
    "sawyer_assembly_obstacle": dict(nq=34, nv=33, nu=7, njnt=28, nbody=53, ngeom=93, collidable=39, by_type={0: 1, 6: 19, 5: 12, 2: 3, 3: 4}),
    "pusher_obstacle": dict(nq=16, nv=16, nu=4, njnt=16, nbody=26, ngeom=33, collidable=16, by_type={3: 7, 6: 8, 5: 1}),
}


@pytest.mark.parametrize("name", sorted(APPENDIX_A))
def test_compiled_sizes(name):
    m, exp = load_model(name), APPENDIX_A[name]
    for k in ("nq", "nv", "nu", "njnt", "nbody", "ngeom"):
        assert getattr(m, k) == exp[k], k
    col = (m.geom_contype | m.geom_conaffinity) != 0
    assert col.sum() == exp["collidable"]
    for t, c in exp["by_type"].items():
        assert (m.geom_type[col] == t).sum() == c, t


def test_push_model_details(push_model):
    m = push_model
    # qpos layout: arm, gripper, two ghost chains, cube free joint, target sliders
    assert [m.get_joint_qpos_addr("right_j%d" % i) for i in range(7)] == list(range(7))
    assert m.get_joint_qpos_addr("cube") == (27, 34) and m.get_joint_qpos_addr("target_y") == 35
    assert np.allclose(m.qpos0[27:34], [0.92, 0, 0.88, 1, 0, 0, 0])
    # body offsets straight from the XML (sawyer_no_gripper_chain.xml:62)
    b = m.body_name2id("right_l1")
    assert np.allclose(m.body_pos[b], [0.081, 0.05, 0.237]) and np.allclose(m.body_quat[b], [0.5, -0.5, 0.5, 0.5])
    # joint ranges / damping / armature (class sawyer defaults + per-joint overrides)
    j = m.joint_name2id("right_j1")
    assert np.allclose(m.jnt_range[j], [-3.8, 1.25]) and m.jnt_limited[j] == 1
    assert m.dof_damping[m.jnt_dofadr[j]] == 50 and m.dof_armature[m.jnt_dofadr[j]] == 0.1
    assert m.dof_damping[m.jnt_dofadr[m.joint_name2id("right_j4")]] == 10
    g = m.joint_name2id("rc_close")                      # childclass sawyer_gripper
    assert m.dof_damping[m.jnt_dofadr[g]] == 100 and m.dof_armature[m.jnt_dofadr[g]] == 5
    # position actuators (sawyer_joint_pos_act.xml:11-17)
    assert np.allclose(m.actuator_kp, [500, 500, 200, 200, 50, 50, 50])
    assert np.allclose(m.actuator_forcerange[:, 1], [100, 100, 75, 75, 50, 50, 50]) and m.actuator_forcelimited.all()
    # geom defaults: sawyer class margin / solref, gripper class friction
    l2 = [g for g in range(m.ngeom) if m.geom_bodyid[g] == m.body_name2id("right_l2") and m.geom_type[g] == 3][0]
    assert m.geom_margin[l2] == 0.001 and np.allclose(m.geom_solref[l2], [0.008, 1]) and np.allclose(m.geom_solimp[l2][:3], [0.95, 0.95, 0.01])
    claw = m.geom_name2id("rightclaw_it")
    assert np.allclose(m.geom_friction[claw], [1, 0.5, 0.001]) and m.geom_condim[claw] == 6
    cube = m.geom_name2id("cube")
    assert m.geom_condim[cube] == 4 and np.isclose(m.body_mass[m.body_name2id("cube")], 300 * 0.06 ** 3)
    # ghost chains never collide
    for g in range(m.ngeom):
        if "indicator" in m.names["body"][m.geom_bodyid[g]] or "_target" in m.names["body"][m.geom_bodyid[g]]:
            assert m.geom_contype[g] == 0 and m.geom_conaffinity[g] == 0
    # unnormalised quat in the XML (bin1 quat="0 1 0 1") is normalised
    assert np.isclose(np.linalg.norm(m.body_quat[m.body_name2id("bin1")]), 1.0)
    assert m.opt_timestep == 0.002 and m.opt_cone == 1 and m.opt_iterations == 50 and m.opt_noslip_iterations == 5
    assert tuple(m.exclude_body[0]) == (m.body_name2id("right_arm_base_link"), m.body_name2id("right_l0"))
    # inertia inferred from geoms where <inertial> is commented out (l4..l6), explicit elsewhere
    assert np.isclose(m.body_mass[m.body_name2id("right_l3")], 2.5097)
    assert 1.0 < m.body_mass[m.body_name2id("right_l4")] < 10.0
    assert (m.body_inertia[m.body_name2id("right_l5")] > 0).all()


@pytest.mark.skipif(not os.path.isdir(REF_XML), reason="reference checkout not present (GPU box)")
def test_cached_models_match_a_fresh_compile():
    from mopa_rl_b200.mjcf import compile_mjcf

    for name in APPENDIX_A:
        fresh, cached = compile_mjcf(os.path.join(REF_XML, name + ".xml")), load_model(name)
        for k, v in fresh.__dict__.items():
            if isinstance(v, np.ndarray) and v.dtype.kind in "fiu":
                assert np.array_equal(v, getattr(cached, k)), (name, k)
