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


def lift_random_qpos(model, n, seed, ref, fk_scene=None, floating=0.5):
    """SawyerLiftObstacle-v0 states: arm ~ U(joint range) and the can ("cube" free joint) either at its keyframe
    pose in the bin or — for a `floating` fraction of the states — at a random orientation near the gripper
    (grip_site + U(-0.1, 0.1)^3, through the oracle's FK when `fk_scene` is given) or anywhere in the arm's workspace,
    so that the mesh collider decides.  All values fp32-representable."""
    rng = np.random.Generator(np.random.PCG64(seed))
    q = random_qpos(model, n, seed + 1, ref)
    a = model.get_joint_qpos_addr("cube")[0]
    fl = rng.random(n) < floating
    pos = np.stack([rng.uniform(0.3, 0.9, n), rng.uniform(-0.5, 0.5, n), rng.uniform(0.8, 1.4, n)], 1)
    if fk_scene is not None:
        site = model.names["site"].index("grip_site")
        off = rng.uniform(-0.1, 0.1, (n, 3))
        for i in np.nonzero(fl)[0][::2]:
            pos[i] = fk_scene.fk(q[i])["site_xpos"][site] + off[i]
    quat = rng.normal(size=(n, 4))
    quat /= np.linalg.norm(quat, axis=1, keepdims=True)
    q[fl, a:a + 3] = pos[fl]
    q[fl, a + 3:a + 7] = quat[fl]
    return q.astype(np.float32).astype(np.float64)
