"""GPU <-> oracle parity of the RRT-Connect kernel: status, waypoints, tree-node indices."""
import numpy as np
import pytest

from helpers import PUSH_INIT_QPOS, planner_setup, random_qpos

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planners(push_model, oracle_built):
    from mopa_rl_b200.capi import NativePlanner

    ignored, passive, ref = planner_setup(push_model)
    native = NativePlanner(push_model, passive, ignored, -0.002, 0.1, seed=1234)
    scene = oracle_built.OracleScene(push_model, ignored, -0.002, "f32")
    adr, lo, hi, so2 = oracle_built.space_from_model(push_model, passive)
    orc = oracle_built.OraclePlanner(scene, adr, lo, hi, so2, 0.1, 0.005, seed=1234, max_nodes=4096)
    return native, scene, orc, ref


def _problems(model, scene, ref, n, seed):
    q = random_qpos(model, 6 * n, seed, ref, spread=0.5)
    v = q[(scene.is_valid(q) & 1) == 1]
    assert len(v) >= 2 * n
    return v[:n], v[n:2 * n]


def test_plan_matches_oracle(planners, push_model):
    native, scene, orc, ref = planners
    n = 96
    start, goal = _problems(push_model, scene, ref, n, 21)
    keys = np.arange(n, dtype=np.uint64) + 1000
    out = native.plan_host(start, goal, keys, max_iter=400, max_path=512)
    n_ok = 0
    for i in range(n):
        r = orc.plan(start[i], goal[i], int(keys[i]), 400, 512)
        assert out["status"][i] == r["status"], i
        assert out["iters"][i] == r["iters"], i
        L = len(r["path"])
        assert out["path_len"][i] == L
        if r["status"] == 0:
            n_ok += 1
            assert np.array_equal(out["node_ids"][i, :L], r["node_ids"]), "waypoint indices differ for problem %d" % i
            assert np.array_equal(out["path"][i, :L], r["path"]), "waypoints differ for problem %d" % i
            # path invariants (OMPL): starts at start, ends at goal, hops within range (L1), every vertex valid
            p = r["path"][:, ref]
            assert np.allclose(p[0], start[i][ref].astype(np.float32)) and np.allclose(p[-1], goal[i][ref].astype(np.float32))
            assert np.abs(np.diff(p, axis=0)).sum(1).max() <= 0.1 + 1e-5
            assert (scene.is_valid(r["path"]) & 1).all()
    assert n_ok >= n // 2


def test_sentinels(planners, push_model):
    native, scene, orc, ref = planners
    q0 = push_model.qpos0.copy()
    q0[ref] = PUSH_INIT_QPOS
    cand = random_qpos(push_model, 64, 5, ref)
    bad = cand[(scene.is_valid(cand) & 1) == 0][0]  # some colliding arm configuration
    assert scene.is_valid(bad)[0] & 1 == 0 and scene.is_valid(q0)[0] & 1 == 1
    out = native.plan_host(np.stack([q0, bad, q0]), np.stack([bad, q0, q0]), [1, 2, 3], max_iter=50)
    assert out["status"][0] == -5 and out["path_len"][0] == 0     # invalid goal -> the reference's -5 row
    assert out["status"][1] == -4 and out["path_len"][1] == 0     # invalid start -> no exact solution (-4 row)
    assert out["status"][2] == 0                                  # start == goal: trivially connected
    oob = q0.copy()
    oob[ref[0]] = 3.2  # outside jnt_range of right_j0
    out = native.plan_host(q0[None], oob[None], [9], max_iter=50)
    assert out["status"][0] in (-4, -5)
    r = orc.plan(q0, oob, 9, 50)
    assert r["status"] == out["status"][0]


def test_passive_dims_frozen_and_key_dependence(planners, push_model):
    native, scene, orc, ref = planners
    start, goal = _problems(push_model, scene, ref, 8, 33)
    a = native.plan_host(start, goal, np.arange(8) + 1, max_iter=300)
    b = native.plan_host(start, goal, np.arange(8) + 1, max_iter=300)
    assert np.array_equal(a["path"], b["path"]) and np.array_equal(a["status"], b["status"])  # deterministic
    passive = [i for i in range(push_model.nq) if i not in ref]
    for i in range(8):
        L = a["path_len"][i]
        if L:
            assert np.array_equal(a["path"][i, :L][:, passive], np.tile(start[i][passive].astype(np.float32), (L, 1)))


def test_lift_plan_matches_oracle(oracle_built):
    """RRT-Connect on the lift scene (mesh collider): the can floats next to the arm so that edges are checked
    against the hull; status / iterations / node ids / waypoints bit-identical to the oracle."""
    from mopa_rl_b200.capi import NativePlanner
    from mopa_rl_b200.model import load_model

    m = load_model("SawyerLiftObstacle-v0")
    ignored, passive, ref = planner_setup(m)
    native = NativePlanner(m, passive, ignored, -0.002, 0.1, seed=77)
    scene = oracle_built.OracleScene(m, ignored, -0.002, "f32")
    adr, lo, hi, so2 = oracle_built.space_from_model(m, passive)
    orc = oracle_built.OraclePlanner(scene, adr, lo, hi, so2, 0.1, 0.005, seed=77, max_nodes=4096)
    n = 32
    q = random_qpos(m, 12 * n, 3, ref, spread=1.0)
    a = m.get_joint_qpos_addr("cube")[0]
    q[:, a:a + 7] = np.array([0.55, 0.1, 1.05, 0.8, 0.0, 0.6, 0.0], dtype=np.float32)   # can in the workspace, tilted
    v = q[(scene.is_valid(q) & 1) == 1]
    assert len(v) >= 2 * n
    start, goal = v[:n], v[n:2 * n]
    keys = np.arange(n, dtype=np.uint64) + 500
    out = native.plan_host(start, goal, keys, max_iter=300, max_path=512)
    n_ok = 0
    for i in range(n):
        r = orc.plan(start[i], goal[i], int(keys[i]), 300, 512)
        assert out["status"][i] == r["status"] and out["iters"][i] == r["iters"], i
        L = len(r["path"])
        assert out["path_len"][i] == L
        if r["status"] == 0:
            n_ok += 1
            assert np.array_equal(out["node_ids"][i, :L], r["node_ids"]) and np.array_equal(out["path"][i, :L], r["path"])
    assert n_ok >= n // 4


def _pusher_setup(m):
    """PusherObstacle-v0 planner inputs as rl/trainer.py:62-75 derives them from env/pusher/pusher_obstacle.py:
    manipulation geom `box` x static obstacle geoms ignored, joints 0-3 active (joint0 unlimited -> SO(2))."""
    static = [m.geom_name2id("obstacle%d_geom" % i) for i in range(1, 8)]
    box = m.geom_name2id("box")
    ignored = [(min(box, g), max(box, g)) for g in static]
    ref = [m.get_joint_qpos_addr("joint%d" % i) for i in range(4)]
    passive = [i for i in range(m.nq) if i not in ref]
    return ignored, passive, ref


def _pusher_states(m, ref, n, seed, spread=1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    q = np.tile(m.qpos0, (n, 1))
    q[:, ref[0]] = rng.uniform(-3.14, 3.14, n)          # unlimited hinge: the reference's +-3.14 convention
    for k in (1, 2, 3):
        j = list(m.jnt_qposadr).index(ref[k])
        q[:, ref[k]] = rng.uniform(m.jnt_range[j, 0] * spread, m.jnt_range[j, 1] * spread, n)
    return q.astype(np.float32).astype(np.float64)


def test_pusher_validity_and_plan_match_oracle(oracle_built):
    """BASELINE configs[0] scene (2-D pusher, 4 hinges, joint0 on SO(2)): validity words and RRT-Connect results
    bit-identical to the oracle with the reference's Pusher settings (range 0.2, contact_threshold -0.0015)."""
    from mopa_rl_b200.capi import NativePlanner
    from mopa_rl_b200.model import load_model

    m = load_model("PusherObstacle-v0")
    ignored, passive, ref = _pusher_setup(m)
    native = NativePlanner(m, passive, ignored, -0.0015, 0.2, seed=9)
    scene = oracle_built.OracleScene(m, ignored, -0.0015, "f32")
    g1, g2 = native.pairs()
    o1, o2 = scene.pairs()
    assert native.n_pairs == scene.npair == 80 and np.array_equal(g1, o1) and np.array_equal(g2, o2)
    q = _pusher_states(m, ref, 50000, 5)
    ow = scene.is_valid(q)
    assert np.array_equal(native.is_valid_host(q, flags=1, return_words=True)[1], ow)
    assert 0.05 < (ow & 1).mean() < 0.5
    adr, lo, hi, so2 = oracle_built.space_from_model(m, passive)
    assert list(so2) == [1, 0, 0, 0]
    orc = oracle_built.OraclePlanner(scene, adr, lo, hi, so2, 0.2, 0.005, seed=9, max_nodes=4096)
    v = q[(ow & 1) == 1]
    n = 48
    start, goal = v[:n], v[n:2 * n]
    keys = np.arange(n, dtype=np.uint64) + 100
    out = native.plan_host(start, goal, keys, max_iter=400, max_path=512)
    n_ok = n_long = 0
    for i in range(n):
        r = orc.plan(start[i], goal[i], int(keys[i]), 400, 512)
        assert out["status"][i] == r["status"] and out["iters"][i] == r["iters"], i
        L = len(r["path"])
        assert out["path_len"][i] == L
        if r["status"] == 0:
            n_ok += 1
            n_long += r["iters"] > 1
            assert np.array_equal(out["node_ids"][i, :L], r["node_ids"]) and np.array_equal(out["path"][i, :L], r["path"])
            # SO(2): joint0 stays in [-pi, pi] and hops are measured around the circle
            p = r["path"][:, ref]
            assert np.abs(p[:, 0]).max() <= np.float32(np.pi)
            d0 = np.abs(np.diff(p[:, 0]))
            hop = np.minimum(d0, 2 * np.pi - d0) + np.abs(np.diff(p[:, 1:], axis=0)).sum(1)
            assert hop.max() <= 0.2 + 1e-5
    assert n_ok >= 5 and n_long >= 2   # random 4-hinge states in the cluttered scene: most pairs are not connectable in 400 iterations
