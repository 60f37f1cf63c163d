"""Shared scene / sampling helpers for the tests (mirrors the planner set-up of rl/trainer.py:62-75)."""
import numpy as np

PUSH_INIT_QPOS = np.array([0.000457, -0.114, 0.0321, -0.00712, 0.0303, -0.0302, -0.00994])


def planner_setup(model, static_bodies=("table", "bin1"), manipulation_geoms=("cube",), robot_joints=None):
    """ignored contact pairs and passive joint indices exactly as Trainer.__init__ derives them."""
    robot_joints = robot_joints or ["right_j%d" % i for i in range(7)]
    static_ids = [g for g in range(model.ngeom) if model.names["body"][model.geom_bodyid[g]] in static_bodies]
    ignored = []
    for name in manipulation_geoms:
        mg = model.geom_name2id(name)
        ignored += [(min(mg, g), max(mg, g)) for g in static_ids]
    ref = [model.get_joint_qpos_addr(j) for j in robot_joints]
    passive = [i for i in range(model.nq) if i not in ref]
    return ignored, passive, ref


def random_qpos(model, n, seed, ref, spread=1.0):
    """qpos[ref] ~ U(joint range) (fp32-representable), passive dims at qpos0 — BASELINE config 5."""
    rng = np.random.Generator(np.random.PCG64(seed))
    jid = [list(model.jnt_qposadr).index(a) for a in ref]
    lo, hi = model.jnt_range[jid, 0] * spread, model.jnt_range[jid, 1] * spread
    q = np.tile(model.qpos0, (n, 1))
    q[:, ref] = rng.uniform(lo, hi, (n, len(ref)))
    return q.astype(np.float32).astype(np.float64)
