"""Collision / state-validity oracle: analytic known answers, filters, scene invariants, goldens."""
import os

import numpy as np
import pytest

from helpers import PUSH_INIT_QPOS, planner_setup, random_qpos

I3 = np.eye(3).ravel()
PLANE, SPHERE, CAPSULE, CYL, BOX = 0, 2, 3, 5, 6
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return (np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K).ravel()


@pytest.mark.parametrize("prec,tol", [("f64", 1e-12), ("f32", 2e-6)])
def test_primitive_signed_distances(oracle_built, prec, tol):
    d = lambda *a: oracle_built.primitive_dist(*a, precision=prec)
    z = [0, 0, 0]
    # touching, 2 mm penetration (= the planner's contact_threshold), separated
    for gap in (0.0, -0.002, 0.05):
        assert abs(d(SPHERE, z, I3, [0.1, 0, 0], SPHERE, [0.3 + gap, 0, 0], I3, [0.2, 0, 0]) - gap) < tol
        assert abs(d(PLANE, z, I3, z, SPHERE, [0.3, 0.2, 0.1 + gap], I3, [0.1, 0, 0]) - gap) < tol
        assert abs(d(PLANE, z, I3, z, BOX, [0, 0, 0.3 + gap], I3, [0.1, 0.2, 0.3]) - gap) < tol
        assert abs(d(PLANE, z, I3, z, CAPSULE, [0, 0, 0.25 + gap], I3, [0.05, 0.2, 0]) - gap) < tol          # upright capsule
        assert abs(d(PLANE, z, I3, z, CYL, [0, 0, 0.2 + gap], I3, [0.05, 0.2, 0]) - gap) < tol
        assert abs(d(PLANE, z, I3, z, CYL, [0, 0, 0.05 + gap], rot([1, 0, 0], np.pi / 2), [0.05, 0.2, 0]) - gap) < tol  # lying cylinder
        assert abs(d(SPHERE, [0.2 + 0.1 + gap, 0, 0], I3, [0.1, 0, 0], BOX, z, I3, [0.2, 0.3, 0.4]) - gap) < tol
        assert abs(d(SPHERE, [0, 0.1 + 0.05 + gap, 0.1], I3, [0.1, 0, 0], CYL, z, I3, [0.05, 0.3, 0]) - gap) < tol   # radial
        assert abs(d(SPHERE, [0.01, 0, 0.3 + 0.1 + gap], I3, [0.1, 0, 0], CYL, z, I3, [0.05, 0.3, 0]) - gap) < tol  # cap
        assert abs(d(SPHERE, [0.15 + gap, 0, 0.1], I3, [0.1, 0, 0], CAPSULE, z, I3, [0.05, 0.3, 0]) - gap) < tol
        assert abs(d(CAPSULE, z, I3, [0.05, 0.3, 0], CAPSULE, [0.12 + gap, 0, 0], rot([1, 0, 0], np.pi / 2), [0.07, 0.2, 0]) - gap) < tol
        assert abs(d(BOX, z, I3, [0.1, 0.1, 0.1], BOX, [0.3 + gap, 0.02, 0.03], I3, [0.2, 0.1, 0.1]) - gap) < tol
    # sphere centre inside a box: distance to the nearest face minus the radius
    assert abs(d(SPHERE, [0.15, 0, 0], I3, [0.01, 0, 0], BOX, z, I3, [0.2, 0.3, 0.4]) - (-0.05 - 0.01)) < tol
    # rotated box-box: edge of a 45-degree box pressed 1 mm into a face
    R45 = rot([0, 0, 1], np.pi / 4)
    got = d(BOX, z, I3, [0.1, 0.1, 0.1], BOX, [0.1 + 0.1 * np.sqrt(2) - 0.001, 0, 0], R45, [0.1, 0.1, 0.1])
    assert abs(got + 0.001) < 10 * tol


def test_mpr_pairs_agree_with_geometry(oracle_built):
    d = lambda *a: oracle_built.primitive_dist(*a, precision="f64")
    z = [0, 0, 0]
    Ry = rot([0, 1, 0], np.pi / 2)
    # capsule lying on a box face, 3 mm deep -> depth 3 mm; separated -> no contact reported (big value)
    assert abs(d(CAPSULE, [0, 0, 0.1 + 0.05 - 0.003], Ry, [0.05, 0.2, 0], BOX, z, I3, [0.3, 0.3, 0.1]) + 0.003) < 1e-5
    assert d(CAPSULE, [0, 0, 0.1 + 0.05 + 0.01], Ry, [0.05, 0.2, 0], BOX, z, I3, [0.3, 0.3, 0.1]) > 1.0
    # cylinder standing on a box, 2.5 mm deep; coaxial cylinders end to end 1 mm deep
    assert abs(d(CYL, [0, 0, 0.1 + 0.2 - 0.0025], I3, [0.05, 0.2, 0], BOX, z, I3, [0.3, 0.3, 0.1]) + 0.0025) < 1e-5
    assert abs(d(CYL, z, I3, [0.05, 0.2, 0], CYL, [0.01, 0, 0.2 + 0.1 - 0.001], I3, [0.08, 0.1, 0]) + 0.001) < 1e-5
    # capsule side against cylinder side
    assert abs(d(CAPSULE, [0.05 + 0.08 - 0.004, 0, 0], I3, [0.05, 0.2, 0], CYL, z, I3, [0.08, 0.3, 0]) + 0.004) < 1e-5


@pytest.fixture(scope="module")
def push_scene(push_model, oracle_built):
    ignored, passive, ref = planner_setup(push_model)
    return (oracle_built.OracleScene(push_model, ignored, -0.002, "f32"), oracle_built.OracleScene(push_model, ignored, -0.002, "f64"),
            oracle_built.OracleScene(push_model, [], -0.002, "f64"), ignored, ref)


def test_forward_kinematics_known_answers(push_scene, push_model):
    m = push_model
    f = push_scene[1].fk(m.qpos0)
    xp = lambda n: f["body_xpos"][m.body_name2id(n)]
    assert np.allclose(xp("base"), [0, 0, 0.95]) and np.allclose(xp("right_l0"), [0, 0, 1.03])
    assert np.allclose(xp("right_l1"), [0.081, 0.05, 1.267], atol=1e-12)
    assert np.allclose(xp("cube"), [0.92, 0, 0.88]) and np.allclose(xp("target"), [1.04, 0, 0.85])
    # rotating right_j0 by 90 degrees about z moves l1 from (0.081, 0.05) to (-0.05, 0.081)
    q = m.qpos0.copy()
    q[0] = np.pi / 2
    assert np.allclose(push_scene[1].fk(q)["body_xpos"][m.body_name2id("right_l1")], [-0.05, 0.081, 1.267], atol=1e-12)
    # rotation matrices stay orthonormal down the chain
    q[:7] = [0.3, -0.5, 0.7, 1.1, -0.2, 0.4, 2.0]
    R = push_scene[1].fk(q)["body_xmat"][m.body_name2id("rightclaw")].reshape(3, 3)
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
    # the target slides along world x / y
    q = m.qpos0.copy()
    q[34:36] = [0.01, -0.02]
    assert np.allclose(push_scene[1].fk(q)["body_xpos"][m.body_name2id("target")], [1.05, -0.02, 0.85])


def test_candidate_pair_filters(push_scene, push_model):
    m = push_model
    s32, s64, s_noignore, ignored, ref = push_scene
    g1, g2 = s_noignore.pairs()
    assert len(g1) == 250                      # SURVEY.md Appendix A census for Sawyer-Push
    assert s32.npair == 250 - 9                # cube x {table (5), bin1 (4)} collidable geoms are ignored
    bodies = lambda g: m.names["body"][m.geom_bodyid[g]]
    for a, b in zip(g1, g2):
        wa, wb = m.body_weldid[m.geom_bodyid[a]], m.body_weldid[m.geom_bodyid[b]]
        assert wa != wb                                                     # never two geoms of one weld group (all statics)
        assert (m.geom_contype[a] & m.geom_conaffinity[b]) or (m.geom_contype[b] & m.geom_conaffinity[a])
        assert {bodies(a), bodies(b)} != {"right_arm_base_link", "right_l0"}     # <exclude>
        assert not ("indicator" in bodies(a) or "indicator" in bodies(b))
    pairs = set(zip(g1.tolist(), g2.tolist()))
    l2 = [g for g in range(m.ngeom) if bodies(g) == "right_l2" and m.geom_contype[g]][0]
    l3 = [g for g in range(m.ngeom) if bodies(g) == "right_l3" and m.geom_contype[g]][0]
    l4 = [g for g in range(m.ngeom) if bodies(g) == "right_l4" and m.geom_contype[g]][0]
    assert (l2, l3) not in pairs and (l2, l4) in pairs                      # parent-child filtered, grand-child kept


def test_scene_invariants(push_scene, push_model):
    m = push_model
    s32, s64, s_noignore, ignored, ref = push_scene
    q = m.qpos0.copy()
    q[ref] = PUSH_INIT_QPOS
    assert s32.is_valid(q)[0] == 1 and s64.is_valid(q)[0] == 1               # nominal reset pose is valid
    # the cube resting on the bin floor: ignored pairs keep the state valid, without them it may not be
    rest = q.copy()
    rest[29] = 0.8596
    assert s32.is_valid(rest)[0] == 1
    sunk = q.copy()
    sunk[29] = 0.85                                                           # 1 cm into the bin floor
    assert s32.is_valid(sunk)[0] == 1 and s_noignore.is_valid(sunk)[0] & 1 == 0
    # ghost-arm joints never matter
    Q = random_qpos(m, 2000, 5, ref)
    base = s32.is_valid(Q)
    Q2 = Q.copy()
    Q2[:, 9:27] = np.random.default_rng(0).uniform(-1, 1, (2000, 18)).astype(np.float32)
    assert np.array_equal(s32.is_valid(Q2), base)
    # an arm folded into the pedestal / table is invalid, and the first offending pair is reported
    assert 0.3 < (base & 1).mean() < 0.6
    bad = base[(base & 1) == 0]
    assert ((bad >> 8) >= 1).all() and ((bad >> 8) <= s32.npair).all()


def test_golden_validity_words(push_scene, push_model):
    s32, s64 = push_scene[0], push_scene[1]
    g = np.load(os.path.join(GOLD, "push_validity.npz"))
    q = np.tile(push_model.qpos0, (len(g["active"]), 1))
    q[:, push_scene[4]] = g["active"].astype(np.float64)
    w, d = s32.is_valid(q, True)
    assert np.array_equal(w, g["words_f32"])
    assert np.allclose(d, g["min_dist_f32"], atol=1e-6)
    w64 = s64.is_valid(q) & 1
    assert np.array_equal(w64, g["valid_f64"])
    # fp32 arithmetic flips no boolean on this sample (report, target 0)
    assert int(((w & 1) != w64).sum()) == 0


# ---------------------------------------------------------------------------- mesh collider (lift scene: the can)
@pytest.fixture(scope="module")
def lift_scene(oracle_built):
    from helpers import planner_setup as ps
    from mopa_rl_b200.model import load_model

    m = load_model("SawyerLiftObstacle-v0")
    ignored, passive, ref = ps(m)
    return (m, oracle_built.OracleScene(m, ignored, -0.002, "f32"), oracle_built.OracleScene(m, ignored, -0.002, "f64"),
            oracle_built.OracleScene(m, [], -0.002, "f64"), ref)


def test_mesh_hull_compiled(lift_scene):
    """The can's convex hull (env/assets/objects/meshes/can.stl, 956 triangles): 106 vertices that contain the geom
    origin strictly (MPR's interior point), bounding radius = farthest hull vertex."""
    m = lift_scene[0]
    cube = m.geom_name2id("cube")
    assert m.nmesh == 1 and int(m.mesh_vertnum[0]) == 106 and m.geom_dataid[cube] == 0
    assert (m.geom_dataid >= 0).sum() == 1                                    # visual meshes carry no hull
    v = m.mesh_vert
    assert np.array_equal(v, v.astype(np.float32).astype(np.float64))         # fp32-representable on both sides
    assert np.isclose(m.geom_rbound[cube], np.linalg.norm(v, axis=1).max())
    assert v[:, 2].min() < -0.04 and v[:, 2].max() > 0.039 and np.abs(v[:, :2]).max() < 0.0252
    from scipy.spatial import ConvexHull

    h = ConvexHull(v)
    assert len(h.vertices) == len(v) and (h.equations[:, 3] < 0).all()


def test_mesh_known_answers(lift_scene):
    """plane - hull analytically, box - hull through MPR: the upright can sunk 5 mm into the table top."""
    m, s32, s64, s_all, ref = lift_scene
    cube = m.geom_name2id("cube")
    a = m.get_joint_qpos_addr("cube")[0]
    g1, g2 = s_all.pairs()
    zmin = m.mesh_vert[:, 2].min()
    table = m.geom_name2id("table_collision")
    plane = [g for g in range(m.ngeom) if m.geom_type[g] == 0 and (m.geom_contype[g] or m.geom_conaffinity[g])][0]
    ip = [i for i in range(len(g1)) if {g1[i], g2[i]} == {plane, cube}][0]
    it = [i for i in range(len(g1)) if {g1[i], g2[i]} == {table, cube}][0]
    q = m.qpos0.copy()
    q[a:a + 7] = [0.66, 0.5, 0.81 - zmin - 0.005, 1, 0, 0, 0]               # table top is at z = 0.40 + 0.41
    d = s_all.pair_dists(q)
    plane_z = m.geom_pos[plane][2] + m.body_pos[m.geom_bodyid[plane]][2]
    assert abs(d[ip] - (q[a + 2] + zmin - plane_z)) < 1e-12
    assert abs(d[it] + 0.005) < 2e-5
    q[a + 2] += 0.006                                                         # 1 mm above the table: no penetration
    assert s_all.pair_dists(q)[it] > 1e9
    # tilted by 90 degrees about x: the can lies on its side, lowest point = hull radius in the (y, z) plane
    q[a:a + 7] = [0.66, 0.5, 0.81 + 0.02, np.sqrt(0.5), np.sqrt(0.5), 0, 0]
    d = s_all.pair_dists(q)
    lowest = m.mesh_vert[:, 1].min()   # world z of a vertex = y_local after the rotation
    assert abs(d[it] - (0.02 + lowest)) < 2e-5 and d[it] < 0


def test_lift_scene_invariants_and_golden(lift_scene):
    from helpers import lift_random_qpos

    m, s32, s64, s_all, ref = lift_scene
    cube = m.geom_name2id("cube")
    assert s32.is_valid(m.qpos0)[0] == 1                                      # keyframe: can in the bin, ignored pairs
    assert s_all.is_valid(m.qpos0)[0] & 1 in (0, 1)
    g = np.load(os.path.join(GOLD, "lift_validity.npz"))
    q = g["qpos"].astype(np.float64)
    assert np.array_equal(q, lift_random_qpos(m, len(q), int(g["seed"]), ref, s64))
    w, d = s32.is_valid(q, True)
    assert np.array_equal(w, g["words_f32"])
    assert np.allclose(d, g["min_dist_f32"], atol=1e-6)
    assert np.array_equal(s64.is_valid(q) & 1, g["valid_f64"])
    # the can decides some of the states: first offending pair involves the mesh geom
    g1, g2 = s32.pairs()
    first = (w[(w & 1) == 0] >> 8).astype(int) - 1
    n_mesh = int(((g1[first] == cube) | (g2[first] == cube)).sum())
    assert n_mesh >= 50, n_mesh
    # moving only the can changes validity of states that were valid
    a = m.get_joint_qpos_addr("cube")[0]
    ok = q[(w & 1) == 1][:512].copy()
    ok[:, a:a + 3] = np.float32(0.0)                                         # can inside the pedestal region at the arm's base
    ok[:, a + 2] = np.float32(0.95)
    assert (s32.is_valid(ok) & 1).mean() < 1.0
