"""Build-time reach analysis of the pair table (csrc/scene_build.cu): a candidate pair is dropped only when its bounding spheres can
never touch, so MuJoCo's per-query filter (mj_collideGeoms -> bounding-sphere test, engine_collision_driver.c; the oracle applies the
same filter) would reject it for every state.  Checked against the oracle's FK on random joint configurations, including ones far
outside the joint ranges (hinges are treated as unlimited) and with the passive joints moved."""
import numpy as np
import pytest



@pytest.mark.parametrize("env", ["SawyerPushObstacle-v0", "SawyerLiftObstacle-v0", "SawyerAssemblyObstacle-v0", "PusherObstacle-v0"])
def test_dropped_pairs_never_pass_the_bounding_sphere_filter(env, oracle_built):
    from mopa_rl_b200.capi import scene_pair_table
    from mopa_rl_b200.model import load_model
    from oracle.oracle import OracleScene

    m = load_model(env)
    from mopa_rl_b200 import envs
    from mopa_rl_b200.rollout import env_planner_inputs

    cls = {"SawyerPushObstacle-v0": envs.VecSawyerPushObstacle, "SawyerLiftObstacle-v0": envs.VecSawyerLiftObstacle,
           "SawyerAssemblyObstacle-v0": envs.VecSawyerAssemblyObstacle, "PusherObstacle-v0": envs.VecPusherObstacle}[env]
    ign, passive, _ = env_planner_inputs(cls, m)        # what the rollout runner hands to the planner
    ref = [i for i in range(m.nq) if i not in set(passive)]
    st, kept = scene_pair_table(m, ign, -0.002)
    sc = OracleScene(m, ign, -0.002)
    g1, g2 = sc.pairs()
    assert st["canonical"] == len(g1) == len(kept) and st["kept"] + st["dropped"] <= st["canonical"]
    assert st["kept"] == int(kept.sum()) and st["entries"] % 32 == 0
    dropped = np.flatnonzero(kept == 0)
    rb, gt = np.asarray(m.geom_rbound), np.asarray(m.geom_type)
    margin = np.maximum(np.asarray(m.geom_margin)[g1], np.asarray(m.geom_margin)[g2])
    rng = np.random.default_rng(5)
    n = 400
    q = np.tile(m.qpos0, (n, 1))
    hinge = [int(m.jnt_qposadr[j]) for j in range(m.njnt) if m.jnt_type[j] == 3]
    q[n // 2:, hinge] = rng.uniform(-7.0, 7.0, (n - n // 2, len(hinge)))           # any angle, limits ignored
    worst = np.inf
    for i in range(n):
        gx = sc.fk(q[i])["geom_xpos"]
        gm = sc.fk(q[i])["geom_xmat"]
        for p in dropped:
            a, b = int(g1[p]), int(g2[p])
            if gt[a] == 0 or gt[b] == 0:
                pl, o = (a, b) if gt[a] == 0 else (b, a)
                nrm = gm[pl].reshape(3, 3)[:, 2]
                gap = float(nrm @ (gx[o] - gx[pl])) - rb[o]                         # lowest point of the bounding sphere above the plane
            else:
                gap = float(np.linalg.norm(gx[a] - gx[b])) - (rb[a] + rb[b] + margin[p])
            worst = min(worst, gap)
    print(env, st, "smallest gap of a dropped pair: %.4f" % worst)
    assert worst > 0.0, worst
    if env == "SawyerPushObstacle-v0":
        assert st["dropped"] >= 40, st      # base links, head and screen against the table / bin geoms
