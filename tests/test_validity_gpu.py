"""GPU <-> oracle parity of the state-validity kernel, through the C ABI (ctypes)."""
import numpy as np
import pytest

from helpers import PUSH_INIT_QPOS, planner_setup, random_qpos

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def push_pair(push_model, oracle_built):
    from mopa_rl_b200.capi import NativePlanner

    ignored, passive, ref = planner_setup(push_model)
    native = NativePlanner(push_model, passive, ignored, -0.002, 0.1, seed=1234)
    orc = oracle_built.OracleScene(push_model, ignored, -0.002, "f32")
    return native, orc, ref


def test_pair_list_matches_oracle(push_pair):
    native, orc, _ = push_pair
    g1, g2 = native.pairs()
    o1, o2 = orc.pairs()
    assert native.n_pairs == orc.npair == 241
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2)


def test_init_pose_valid(push_pair, push_model):
    native, orc, ref = push_pair
    q = push_model.qpos0.copy()
    q[ref] = PUSH_INIT_QPOS
    assert native.is_valid_host(q)[0] == 1


@pytest.mark.parametrize("seed,n", [(1234, 200000), (7, 50000)])
def test_random_states_bit_exact(push_pair, push_model, seed, n):
    native, orc, ref = push_pair
    q = random_qpos(push_model, n, seed, ref)
    ow = orc.is_valid(q)
    v, w = native.is_valid_host(q, flags=1, return_words=True)
    assert np.array_equal(w, ow), "first-offending-pair words differ at %d states" % (w != ow).sum()
    vf = native.is_valid_host(q, flags=0)
    assert np.array_equal(vf, (ow & 1).astype(np.uint8))
    assert 0.2 < vf.mean() < 0.7


def test_edge_sizes(push_pair, push_model):
    native, orc, ref = push_pair
    for n in (1, 2, 127, 128, 129, 1000):
        q = random_qpos(push_model, n, 100 + n, ref)
        assert np.array_equal(native.is_valid_host(q, flags=1, return_words=True)[1], orc.is_valid(q))
    assert len(native.is_valid_host(np.zeros((0, push_model.nq)))) == 0
    with pytest.raises(ValueError):
        native.is_valid_host(np.zeros(push_model.nq - 1))


def test_active_joint_states_give_the_same_words_as_full_rows(push_pair, push_model):
    """KinematicPlanner::isValidState's own convention (KinematicPlanner.cpp:253-286): a state is the vector of the planned joints,
    the passive joints come from the planner's qpos.  Sizes on both sides of the 448-query CTA switch, both result formats,
    a moved cube in the base row."""
    native, orc, ref = push_pair
    assert native.n_active == len(ref) == 7
    base = push_model.qpos0.copy()
    a = push_model.get_joint_qpos_addr("cube")[0]
    base[a:a + 3] += [0.03, -0.05, 0.02]
    base = base.astype(np.float32).astype(np.float64)
    for n in (0, 1, 1000, 70000, 150001):
        q = random_qpos(push_model, n, 300 + n, ref) if n else np.zeros((0, push_model.nq))
        q[:, [i for i in range(push_model.nq) if i not in ref]] = base[[i for i in range(push_model.nq) if i not in ref]]
        act = np.ascontiguousarray(q[:, ref], dtype=np.float32)
        ow = orc.is_valid(q) if n else np.zeros(0, np.uint32)
        w = native.is_valid_active(act, base, flags=1) if n else np.zeros(0, np.uint32)
        assert np.array_equal(w, ow), (n, int((w != ow).sum()))
        if n:
            assert np.array_equal(native.is_valid_active(act, base), (ow & 1).astype(bool))
    with pytest.raises(ValueError):
        native.is_valid_active(np.zeros((3, 6), np.float32), base)


def test_passive_dims_matter_only_where_they_should(push_pair, push_model):
    """Ghost-arm joints (contype=conaffinity=0 chains) never change validity; the cube pose does."""
    native, orc, ref = push_pair
    q = random_qpos(push_model, 20000, 99, ref)
    base = native.is_valid_host(q)
    q2 = q.copy()
    rng = np.random.default_rng(5)
    q2[:, 9:27] = rng.uniform(-1, 1, (len(q), 18)).astype(np.float32)
    assert np.array_equal(native.is_valid_host(q2), base)
    q3 = q.copy()
    q3[:, 27:30] = np.float32(0.6), np.float32(0.0), np.float32(1.2)  # cube floating in the arm's workspace
    w3 = native.is_valid_host(q3, flags=1, return_words=True)[1]
    assert np.array_equal(w3, orc.is_valid(q3))
    assert (w3 & 1).mean() < base.mean()


# ---------------------------------------------------------------------------- lift scene: mesh collider (hull of the can)
@pytest.fixture(scope="module")
def lift_pair(oracle_built):
    from mopa_rl_b200.capi import NativePlanner
    from mopa_rl_b200.model import load_model

    m = load_model("SawyerLiftObstacle-v0")
    ignored, passive, ref = planner_setup(m)
    native = NativePlanner(m, passive, ignored, -0.002, 0.1, seed=1234)
    orc = oracle_built.OracleScene(m, ignored, -0.002, "f32")
    orc64 = oracle_built.OracleScene(m, ignored, -0.002, "f64")
    return m, native, orc, orc64, ref, ignored, passive


def test_lift_pair_list_and_golden(lift_pair):
    import os

    m, native, orc, orc64, ref, _, _ = lift_pair
    g1, g2 = native.pairs()
    o1, o2 = orc.pairs()
    assert native.n_pairs == orc.npair and np.array_equal(g1, o1) and np.array_equal(g2, o2)
    cube = m.geom_name2id("cube")
    assert ((g1 == cube) | (g2 == cube)).sum() == 22
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lift_validity.npz"))
    w = native.is_valid_host(g["qpos"].astype(np.float64), flags=1, return_words=True)[1]
    assert np.array_equal(w, g["words_f32"])


@pytest.mark.parametrize("seed,n", [(11, 60000), (12, 20000)])
def test_lift_random_states_bit_exact(lift_pair, seed, n):
    from helpers import lift_random_qpos

    m, native, orc, orc64, ref, _, _ = lift_pair
    q = lift_random_qpos(m, n, seed, ref, orc64 if n <= 20000 else None)
    ow = orc.is_valid(q)
    w = native.is_valid_host(q, flags=1, return_words=True)[1]
    assert np.array_equal(w, ow), "first-offending-pair words differ at %d states" % (w != ow).sum()
    assert np.array_equal(native.is_valid_host(q, flags=0), (ow & 1).astype(np.uint8))
    assert 0.1 < (ow & 1).mean() < 0.7


def test_lift_can_only_scene_bit_exact(lift_pair, oracle_built):
    """Every pair that does not involve the can ignored: validity is decided by the mesh pairs alone
    (plane / sphere / capsule / cylinder / box against the hull)."""
    from helpers import lift_random_qpos
    from mopa_rl_b200.capi import NativePlanner

    m, _, orc, orc64, ref, ignored, passive = lift_pair
    cube = m.geom_name2id("cube")
    g1, g2 = orc.pairs()
    others = [(int(a), int(b)) for a, b in zip(g1, g2) if a != cube and b != cube]
    native = NativePlanner(m, passive, ignored + others, -0.002, 0.1, seed=1)
    o = oracle_built.OracleScene(m, ignored + others, -0.002, "f32")
    assert native.n_pairs == o.npair == 22
    q = lift_random_qpos(m, 16384, 5, ref, orc64, floating=1.0)
    ow = o.is_valid(q)
    w = native.is_valid_host(q, flags=1, return_words=True)[1]
    assert np.array_equal(w, ow), "mesh-pair words differ at %d states" % (w != ow).sum()
    assert 0.02 < ((ow & 1) == 0).mean() < 0.9
