"""RRT-Connect oracle: OMPL invariants, sentinels, golden traces."""
import os

import numpy as np
import pytest

from helpers import PUSH_INIT_QPOS, planner_setup, random_qpos

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def setup(push_model, oracle_built):
    ignored, passive, ref = planner_setup(push_model)
    scene = oracle_built.OracleScene(push_model, ignored, -0.002, "f32")
    adr, lo, hi, so2 = oracle_built.space_from_model(push_model, passive)
    assert adr == ref and so2 == [0] * 7 and np.allclose(lo[1], -3.8) and np.allclose(hi[1], 1.25)
    return scene, oracle_built.OraclePlanner(scene, adr, lo, hi, so2, 0.1, 0.005, seed=1234), ref, passive


def test_golden_traces(setup, push_model):
    scene, pl, ref, passive = setup
    g = np.load(os.path.join(GOLD, "push_rrt.npz"))
    assert (g["status"] == 0).sum() >= 12 and (g["iters"] > 1).any()
    for i in range(len(g["keys"])):
        s, t = push_model.qpos0.copy(), push_model.qpos0.copy()
        s[ref], t[ref] = g["starts"][i], g["goals"][i]
        r = pl.plan(s, t, int(g["keys"][i]), int(g["max_iter"]), 512)
        assert r["status"] == g["status"][i] and r["iters"] == g["iters"][i]
        L = g["path_len"][i]
        assert len(r["path"]) == L
        assert np.array_equal(r["node_ids"], g["node_ids"][i, :L])
        assert np.array_equal(r["path"][:, ref].astype(np.float32), g["paths"][i, :L])


def test_path_invariants(setup, push_model):
    scene, pl, ref, passive = setup
    cand = random_qpos(push_model, 300, 8, ref, spread=0.5)
    v = cand[(scene.is_valid(cand) & 1) == 1]
    solved = 0
    for i in range(10):
        r = pl.plan(v[i], v[10 + i], 50 + i, 300)
        if r["status"] != 0:
            assert r["status"] == -4 and len(r["path"]) == 0
            continue
        solved += 1
        p = r["path"]
        assert np.array_equal(p[0, ref], v[i][ref]) and np.array_equal(p[-1, ref], v[10 + i][ref])   # start .. goal (fp32 inputs)
        assert np.array_equal(p[:, passive], np.tile(v[i][passive], (len(p), 1)))                      # passive dims frozen at start
        hop = np.abs(np.diff(p[:, ref], axis=0)).sum(1)
        assert hop.max() <= 0.1 + 1e-5 and hop.min() > 0                                               # range in the L1 metric
        assert (scene.is_valid(p) & 1).all()                                                           # every vertex valid
        # edges are valid at the checking resolution (0.005 x joint extent)
        mid = 0.5 * (p[1:] + p[:-1])
        assert (scene.is_valid(mid.astype(np.float32).astype(np.float64)) & 1).all()
        ids = r["node_ids"]
        goal_side = (ids >> 30) & 1
        assert goal_side[0] == 0 and goal_side[-1] == 1 and (np.diff(goal_side) >= 0).all()            # start tree then goal tree
        assert ids[0] == 0 and (ids[-1] & ~(1 << 30)) == 0                                             # roots at both ends
    assert solved >= 6
    # different keys -> different sampling sequences; same key -> identical plan
    hard = [i for i in range(10) if pl.plan(v[i], v[10 + i], 1, 300)["iters"] > 1]
    if hard:
        i = hard[0]
        a, b, c = pl.plan(v[i], v[10 + i], 5, 300), pl.plan(v[i], v[10 + i], 5, 300), pl.plan(v[i], v[10 + i], 6, 300)
        assert np.array_equal(a["path"], b["path"])
        assert a["iters"] != c["iters"] or len(a["path"]) != len(c["path"]) or not np.array_equal(a["path"], c["path"])


def test_sentinel_statuses(setup, push_model):
    scene, pl, ref, passive = setup
    q0 = push_model.qpos0.copy()
    q0[ref] = PUSH_INIT_QPOS
    cand = random_qpos(push_model, 64, 5, ref)
    bad = cand[(scene.is_valid(cand) & 1) == 0][0]
    assert pl.plan(q0, bad, 1, 50)["status"] == -5          # invalid goal  -> the reference's -5 row
    assert pl.plan(bad, q0, 1, 50)["status"] == -4          # invalid start -> no exact solution (-4 row)
    r = pl.plan(q0, q0, 1, 50)
    assert r["status"] == 0 and len(r["path"]) >= 1
    oob = q0.copy()
    oob[ref[0]] = 3.2                                         # outside jnt_range of right_j0 (+-3.0503)
    assert pl.plan(q0, oob, 1, 50)["status"] in (-4, -5)
    assert pl.plan(q0, cand[(scene.is_valid(cand) & 1) == 1][0], 3, 0)["status"] == -4   # zero iterations: no solution


def test_pusher_scene_golden_and_so2_invariants(oracle_built):
    """PusherObstacle-v0 (BASELINE configs[0]): joint0 is an unlimited hinge -> SO(2) sub-space
    (mujoco_ompl_interface.cpp:149-281), the other three hinges are bounded; reference Pusher settings
    (config/pusher.py: range 0.2, contact_threshold -0.0015).  Validity words and RRT-Connect traces against the golden
    file; waypoints stay in [-pi, pi] on joint0 and hops are measured around the circle."""
    import os

    from mopa_rl_b200.model import load_model

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pusher_validity_rrt.npz"))
    m = load_model("PusherObstacle-v0")
    static = [m.geom_name2id("obstacle%d_geom" % i) for i in range(1, 8)]
    box = m.geom_name2id("box")
    ignored = [(min(box, s), max(box, s)) for s in static]
    ref = [m.get_joint_qpos_addr("joint%d" % i) for i in range(4)]
    passive = [i for i in range(m.nq) if i not in ref]
    s32 = oracle_built.OracleScene(m, ignored, -0.0015, "f32")
    s64 = oracle_built.OracleScene(m, ignored, -0.0015, "f64")
    assert s32.npair == 80
    q = np.tile(m.qpos0, (len(g["active"]), 1))
    q[:, ref] = g["active"].astype(np.float64)
    w = s32.is_valid(q)
    assert np.array_equal(w, g["words_f32"]) and np.array_equal(s64.is_valid(q) & 1, g["valid_f64"])
    assert int(((w & 1) != g["valid_f64"]).sum()) <= 2                      # fp32 vs fp64 flips (report; 0-2 on 4096 states)
    adr, lo, hi, so2 = oracle_built.space_from_model(m, passive)
    assert list(so2) == [1, 0, 0, 0] and abs(hi[0] - np.pi) < 1e-12 and lo[1] == -3.0
    pl = oracle_built.OraclePlanner(s32, adr, lo, hi, so2, 0.2, 0.005, seed=9)
    v = q[(w & 1) == 1]
    n = int(g["n_plans"])
    n_ok = 0
    for i in range(n):
        r = pl.plan(v[i], v[n + i], int(g["keys"][i]), int(g["max_iter"]), 512)
        L = len(r["path"])
        assert r["status"] == g["status"][i] and r["iters"] == g["iters"][i] and L == g["path_len"][i]
        if r["status"] == 0:
            n_ok += 1
            assert np.array_equal(r["node_ids"], g["node_ids"][i, :L])
            p = r["path"][:, ref]
            assert np.array_equal(p.astype(np.float32), g["paths"][i, :L])
            assert np.abs(p[:, 0]).max() <= np.float32(np.pi)
            d0 = np.abs(np.diff(p[:, 0]))
            hop = np.minimum(d0, 2 * np.pi - d0) + np.abs(np.diff(p[:, 1:], axis=0)).sum(1)
            assert len(hop) == 0 or hop.max() <= 0.2 + 1e-5
            assert (s32.is_valid(r["path"]) & 1).all()
    assert n_ok >= 3
