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
