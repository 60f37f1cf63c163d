"""Generates the self-consistency golden fixtures under tests/golden/ from the CPU oracle.

The reference has no tests or golden vectors of its own and cannot be executed
here, so these pin the ORACLE (and through it the CUDA path) against accidental change; they are
not MuJoCo / OMPL ground truth.  Re-run after an intentional change of the restated algorithms:

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PUSH_INIT_QPOS, lift_random_qpos, planner_setup, random_qpos  # noqa: E402
from mopa_rl_b200.dynmodel import DynModel  # noqa: E402
from mopa_rl_b200.envs import push_reset_state  # noqa: E402
from mopa_rl_b200.model import load_model  # noqa: E402
from oracle import oracle  # noqa: E402
from oracle.env_oracle import PushEnvOracle  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
oracle.build()
m = load_model("SawyerPushObstacle-v0")
ignored, passive, ref = planner_setup(m)
scene = oracle.OracleScene(m, ignored, -0.002, "f32")
scene64 = oracle.OracleScene(m, ignored, -0.002, "f64")

# 1. state validity: seeded qpos -> result words (f32 oracle) and booleans of the f64 build
q = random_qpos(m, 4096, 1234, ref)
w32, d32 = scene.is_valid(q, True)
w64 = scene64.is_valid(q)
np.savez_compressed(os.path.join(out, "push_validity.npz"), seed=1234, active=q[:, ref].astype(np.float32), words_f32=w32,
                    valid_f64=(w64 & 1).astype(np.uint8), min_dist_f32=d32.astype(np.float32))

# 2. RRT-Connect traces: seeded (start, goal) -> status, iterations, node ids, waypoints
adr, lo, hi, so2 = oracle.space_from_model(m, passive)
pl = oracle.OraclePlanner(scene, adr, lo, hi, so2, 0.1, 0.005, seed=1234)
cand = random_qpos(m, 400, 21, ref, spread=0.5)
v = cand[(scene.is_valid(cand) & 1) == 1]
starts, goals, status, iters, plen, ids, paths = [], [], [], [], [], [], []
for i in range(16):
    r = pl.plan(v[i], v[16 + i], 1000 + i, 400, 512)
    starts.append(v[i][ref]), goals.append(v[16 + i][ref]), status.append(r["status"]), iters.append(r["iters"])
    plen.append(len(r["path"]))
    pad_ids = np.full(512, -1, np.int32)
    pad_ids[: len(r["node_ids"])] = r["node_ids"]
    ids.append(pad_ids)
    pp = np.zeros((512, 7), np.float32)
    pp[: len(r["path"])] = r["path"][:, ref]
    paths.append(pp)
np.savez_compressed(os.path.join(out, "push_rrt.npz"), starts=np.array(starts, np.float32), goals=np.array(goals, np.float32),
                    keys=np.arange(16) + 1000, max_iter=400, status=np.array(status), iters=np.array(iters), path_len=np.array(plen),
                    node_ids=np.array(ids), paths=np.array(paths))

# 3. env.step trajectories: seeded resets + action table -> qpos / qvel / reward after each of 3 env steps
dm = DynModel(m)
n = 4
q0, v0 = push_reset_state(m, 77, np.arange(n), np.zeros(n, dtype=np.int64))
rng = np.random.default_rng(5)
acts = rng.uniform(-1, 1, (3, n, 7)).astype(np.float32)
Q, V, R = np.zeros((3, n, m.nq)), np.zeros((3, n, m.nv)), np.zeros((3, n))
for e in range(n):
    env = PushEnvOracle(m, dm)
    env.reset_to(q0[e], v0[e])
    for s in range(3):
        _, R[s, e], _ = env.step(acts[s, e].astype(np.float64))
        Q[s, e], V[s, e] = env.qpos, env.qvel
np.savez_compressed(os.path.join(out, "push_env_steps.npz"), seed=77, actions=acts, qpos=Q, qvel=V, reward=R)

# 4. assembly env.step trajectories (SawyerAssemblyObstacle-v0): qpos / qvel / reward / obs after each of 2 env steps
from mopa_rl_b200.envs import assembly_reset_state  # noqa: E402
from oracle.env_oracle import AssemblyEnvOracle  # noqa: E402

ma = load_model("SawyerAssemblyObstacle-v0")
dma = DynModel(ma)
na = 2
qa, va = assembly_reset_state(ma, 31, np.arange(na), np.zeros(na, dtype=np.int64))
acts_a = np.random.default_rng(9).uniform(-1, 1, (2, na, 7)).astype(np.float32)
Qa, Va, Ra, Oa = np.zeros((2, na, ma.nq)), np.zeros((2, na, ma.nv)), np.zeros((2, na)), np.zeros((2, na, 38))
for e in range(na):
    env = AssemblyEnvOracle(ma, dma)
    env.reset_to(qa[e], va[e])
    for s_ in range(2):
        Oa[s_, e], Ra[s_, e], _ = env.step(acts_a[s_, e].astype(np.float64))
        Qa[s_, e], Va[s_, e] = env.qpos, env.qvel
np.savez_compressed(os.path.join(out, "assembly_env_steps.npz"), seed=31, actions=acts_a, qpos=Qa, qvel=Va, reward=Ra, obs=Oa)

# 5. inverse kinematics: seeded (start, target) -> qpos, error, steps, success (position-only and full pose)
from mopa_rl_b200.envs import make_push_task  # noqa: E402
from mopa_rl_b200.inverse_kinematics import site_frame  # noqa: E402
from oracle.ik_oracle import IKOracle, _mat2quat  # noqa: E402

task = make_push_task(m, dm)
body, local = site_frame(m, dm, "grip_site")
ik = IKOracle(dm, body, local, [int(task.arm_dof[k]) for k in range(7)])
rng = np.random.default_rng(3)
nk = 8
q0 = np.tile(m.qpos0, (nk, 1))
q0[:, :7] = PUSH_INIT_QPOS + rng.normal(0, 0.02, (nk, 7))
tp, tq, res_p, res_q = np.zeros((nk, 3)), np.zeros((nk, 4)), [], []
for i in range(nk):
    qt = q0[i].copy()
    qt[:7] += rng.uniform(-0.4, 0.4, 7)
    sp, R, _ = ik.site_pose(qt)
    tp[i], tq[i] = sp + (rng.uniform(-1.5, 1.5, 3) if i % 4 == 3 else 0.0), _mat2quat(R)
    res_p.append(ik.solve(q0[i], tp[i], None, tol=1e-2))
    res_q.append(ik.solve(q0[i], tp[i], tq[i], tol=1e-2))
np.savez_compressed(os.path.join(out, "push_ik.npz"), q0=q0, target_pos=tp, target_quat=tq,
                    qpos_p=np.array([r[0] for r in res_p]), err_p=np.array([r[1] for r in res_p]), steps_p=np.array([r[2] for r in res_p]),
                    ok_p=np.array([r[3] for r in res_p]), qpos_q=np.array([r[0] for r in res_q]), err_q=np.array([r[1] for r in res_q]),
                    steps_q=np.array([r[2] for r in res_q]), ok_q=np.array([r[3] for r in res_q]))

# 6. scalar rollout loop with reuse_data: per-record checksum columns of 10 macro actions of one environment
from mopa_rl_b200 import rng as crng  # noqa: E402
from mopa_rl_b200.rollout import MoPAConfig, planner_inputs  # noqa: E402
from oracle.rollout_oracle import ScalarMoPARunner  # noqa: E402


def _policy(g, k):
    u = crng.uniform01(3, np.uint64(g), np.uint64(k), np.arange(7, dtype=np.uint64))
    return (2.0 * u - 1.0).astype(np.float32)


ign2, pas2, _ = planner_inputs(m)
runner = ScalarMoPARunner(m, dm, MoPAConfig(max_iter=150, seed=17, reuse_data=True), ign2, pas2, 7, 2024, _policy, max_episode_steps=30)
recs = []
for _ in range(10):
    recs.append(runner.macro_step())
    recs.extend(runner.extra_records)
np.savez_compressed(os.path.join(out, "push_rollout_reuse.npz"), records=np.array(recs, np.float32))

# 7. lift scene (mesh collider: convex hull of the can): validity words for seeded states (arm + can pose)
ml = load_model("SawyerLiftObstacle-v0")
ign_l, passive_l, ref_l = planner_setup(ml)
sl32 = oracle.OracleScene(ml, ign_l, -0.002, "f32")
sl64 = oracle.OracleScene(ml, ign_l, -0.002, "f64")
ql = lift_random_qpos(ml, 4096, 4321, ref_l, sl64)
wl, dl = sl32.is_valid(ql, True)
np.savez_compressed(os.path.join(out, "lift_validity.npz"), seed=4321, qpos=ql.astype(np.float32), words_f32=wl,
                    valid_f64=(sl64.is_valid(ql) & 1).astype(np.uint8), min_dist_f32=dl.astype(np.float32))

# 8. lift env (SawyerLiftObstacle-v0): open the gripper, place the can between the fingers, close, lift - grasp / lift /
#    success rewards, contact list and the 8-D action with the gripper entry
from mopa_rl_b200.envs import lift_reset_state  # noqa: E402
from oracle.env_oracle import LiftEnvOracle, _q2m  # noqa: E402
from mopa_rl_b200.mjcf import mat_to_quat  # noqa: E402


def lift_can_between_fingers(model, dm, e):
    tips = []
    for name in ("l_fingertip_g0", "r_fingertip_g0"):
        g = model.geom_name2id(name)
        sb = dm.bodies.index(int(model.geom_bodyid[g]))
        tips.append(e.xpos[sb] + _q2m(e.xquat[sb]) @ model.geom_pos[g])
    mid, cdir = 0.5 * (tips[0] + tips[1]), (tips[0] - tips[1]) / np.linalg.norm(tips[0] - tips[1])
    R = _q2m(e.xquat[e.b_ee])
    ax = R[:, 1] - (R[:, 1] @ cdir) * cdir
    ax /= np.linalg.norm(ax)
    return mid, mat_to_quat(np.stack([cdir, np.cross(ax, cdir), ax], 1))


dml = DynModel(ml)
nl = 2
ql0, vl0 = lift_reset_state(ml, 13, np.arange(nl), np.zeros(nl, dtype=np.int64))
ql0[1, ml.get_joint_qpos_addr("right_j1")] = -0.6           # second env: fingertips at the success height
rngl = np.random.default_rng(4)
acts_l = np.zeros((7, nl, 8), np.float32)
acts_l[:2, :, :7] = rngl.uniform(-0.2, 0.2, (2, nl, 7))
acts_l[:2, :, 7] = -1.0                                      # open
acts_l[2:, :, 7] = 0.004                                     # close gently
acts_l[4:, :, :7] = rngl.uniform(-1, 1, (3, nl, 7))          # move the arm with the can in hand
Ql, Vl, Rl, Ol, Gl = np.zeros((7, nl, ml.nq)), np.zeros((7, nl, ml.nv)), np.zeros((7, nl)), np.zeros((7, nl, 35)), np.zeros((7, nl), np.uint8)
ca, cva = ml.get_joint_qpos_addr("cube")[0], ml.get_joint_qvel_addr("cube")[0]
for e in range(nl):
    env = LiftEnvOracle(ml, dml, max_episode_steps=50)
    env.reset_to(ql0[e], vl0[e])
    for s in range(7):
        if s == 2:
            mid, quat = lift_can_between_fingers(ml, dml, env)
            q, v = env.qpos.copy(), env.qvel.copy()
            q[ca:ca + 3], q[ca + 3:ca + 7], v[cva:cva + 6] = mid, quat, 0.0
            env.set_state(q, v)
            env.prev_state = None
        if env.terminal:
            break
        Ol[s, e], Rl[s, e], _ = env.step(acts_l[s, e].astype(np.float64))
        Ql[s, e], Vl[s, e], Gl[s, e] = env.qpos, env.qvel, env.has_grasp
np.savez_compressed(os.path.join(out, "lift_env_steps.npz"), seed=13, actions=acts_l, qpos=Ql, qvel=Vl, reward=Rl, obs=Ol, grasp=Gl)
print("lift env golden: rewards", Rl.round(3).tolist(), "grasp", Gl.tolist())

# 9. Pusher scene (BASELINE configs[0]: 2-D pusher, joint0 on SO(2), range 0.2, contact_threshold -0.0015):
#    validity words and RRT-Connect traces
mp = load_model("PusherObstacle-v0")
static_p = [mp.geom_name2id("obstacle%d_geom" % i) for i in range(1, 8)]
box_p = mp.geom_name2id("box")
ign_p = [(min(box_p, g), max(box_p, g)) for g in static_p]
ref_p = [mp.get_joint_qpos_addr("joint%d" % i) for i in range(4)]
pas_p = [i for i in range(mp.nq) if i not in ref_p]
sp32 = oracle.OracleScene(mp, ign_p, -0.0015, "f32")
sp64 = oracle.OracleScene(mp, ign_p, -0.0015, "f64")
rngp = np.random.Generator(np.random.PCG64(5))
qp = np.tile(mp.qpos0, (4096, 1))
qp[:, ref_p[0]] = rngp.uniform(-3.14, 3.14, 4096)
for k in (1, 2, 3):
    jp = list(mp.jnt_qposadr).index(ref_p[k])
    qp[:, ref_p[k]] = rngp.uniform(mp.jnt_range[jp, 0], mp.jnt_range[jp, 1], 4096)
qp = qp.astype(np.float32).astype(np.float64)
wp = sp32.is_valid(qp)
adr_p, lo_p, hi_p, so2_p = oracle.space_from_model(mp, pas_p)
plp = oracle.OraclePlanner(sp32, adr_p, lo_p, hi_p, so2_p, 0.2, 0.005, seed=9)
vp = qp[(wp & 1) == 1]
st_p, it_p, len_p, ids_p, paths_p = [], [], [], [], []
for i in range(24):
    r = plp.plan(vp[i], vp[24 + i], 100 + i, 400, 512)
    st_p.append(r["status"]), it_p.append(r["iters"]), len_p.append(len(r["path"]))
    pid = np.full(512, -1, np.int32)
    pid[: len(r["node_ids"])] = r["node_ids"]
    ids_p.append(pid)
    pp = np.zeros((512, 4), np.float32)
    pp[: len(r["path"])] = r["path"][:, ref_p]
    paths_p.append(pp)
np.savez_compressed(os.path.join(out, "pusher_validity_rrt.npz"), active=qp[:, ref_p].astype(np.float32), words_f32=wp,
                    valid_f64=(sp64.is_valid(qp) & 1).astype(np.uint8), n_plans=24, keys=np.arange(24) + 100, max_iter=400,
                    status=np.array(st_p), iters=np.array(it_p), path_len=np.array(len_p), node_ids=np.array(ids_p), paths=np.array(paths_p))
print("pusher golden: valid fraction %.3f, plan status %s" % ((wp & 1).mean(), st_p))

# 10. scalar rollout loop, discrete_action (omega = 0, ac_type picks the branch) and the lift task (8-D actions): records of one env
cfg_d = MoPAConfig(max_iter=150, seed=23, omega=0.0, discrete_action=True, reuse_data=True, max_reuse_data=15)


def _policy_d(g, k):
    u = crng.uniform01(13, np.uint64(g), np.uint64(k), np.arange(8, dtype=np.uint64))
    return (2.0 * u[:7] - 1.0).astype(np.float32), bool(u[7] < 0.5)


run_d = ScalarMoPARunner(m, dm, cfg_d, ign2, pas2, 300, 606, _policy_d, max_episode_steps=30)
recs_d = []
for _ in range(12):
    recs_d.append(run_d.macro_step())
    recs_d.extend(run_d.extra_records)
from mopa_rl_b200.envs import VecSawyerLiftObstacle  # noqa: E402
from mopa_rl_b200.rollout import env_planner_inputs  # noqa: E402

cfg_l = MoPAConfig(max_iter=150, seed=31, reuse_data=True, max_reuse_data=15)


def _policy_l(g, k):
    u = crng.uniform01(19, np.uint64(g), np.uint64(k), np.arange(8, dtype=np.uint64))
    return (2.0 * u - 1.0).astype(np.float32)


ign_l2, pas_l2, _ = env_planner_inputs(VecSawyerLiftObstacle, ml)
run_l = ScalarMoPARunner(ml, DynModel(ml), cfg_l, ign_l2, pas_l2, 60, 515, _policy_l, max_episode_steps=20, task="lift")
recs_l = []
for _ in range(8):
    recs_l.append(run_l.macro_step())
    recs_l.extend(run_l.extra_records)
np.savez_compressed(os.path.join(out, "rollout_discrete_lift.npz"), discrete=np.array(recs_d, np.float32), lift=np.array(recs_l, np.float32))
print("rollout goldens: discrete %d records, lift %d records" % (len(recs_d), len(recs_l)))

# 11. Pusher env (RK4 + PID + velocity actuators): the stretched arm sweeps into the box
from oracle.env_oracle import PusherEnvOracle  # noqa: E402

dmp = DynModel(mp)
envp = PusherEnvOracle(mp, dmp)
qp0 = mp.qpos0.copy()
qp0[-2:], qp0[-4:-2] = [0.36, 0.06], [-0.3, 0.15]
envp.reset_to(qp0, np.zeros(mp.nv))
acts_p = np.tile(np.array([0.1, 0.0, 0.0, 0.0]), (4, 1))
acts_p[2:, 1] = -0.05
Qp, Vp, Rp, Op, Np = [], [], [], [], []
for s in range(4):
    ob, r, _ = envp.step(acts_p[s])
    Qp.append(envp.qpos.copy()), Vp.append(envp.qvel.copy()), Rp.append(r), Op.append(ob), Np.append(envp.ncon)
np.savez_compressed(os.path.join(out, "pusher_env_steps.npz"), qpos0=qp0, actions=acts_p, qpos=np.array(Qp), qvel=np.array(Vp),
                    reward=np.array(Rp), obs=np.array(Op), ncon=np.array(Np))
print("pusher env golden: box", Qp[-1][-2:], "ncon", Np, "rewards", np.round(Rp, 4))

# 12. scalar MoPA loop on the Pusher (BASELINE configs[0], scripts/2d/mopa.sh: omega 0.5, action_range 1.0, reuse_data)
ign_pp = [(min(box_p, g), max(box_p, g)) for g in static_p]
cfg_p = MoPAConfig(omega=0.5, action_range=1.0, ac_scale=0.1, step_size=0.04, joint_margin=0.0, contact_threshold=-0.0015, range=0.2,
                   max_iter=1000, reuse_data=True, max_reuse_data=30, seed=5)


def _policy_p(g, k):
    u = crng.uniform01(3, np.uint64(g), np.uint64(k), np.arange(4, dtype=np.uint64))
    return (2.0 * u - 1.0).astype(np.float32)


run_p = ScalarMoPARunner(mp, DynModel(mp), cfg_p, ign_pp, pas_p, 0, 11, _policy_p, max_episode_steps=400, task="pusher")
recs_p = []
for _ in range(12):
    recs_p.append(run_p.macro_step())
    recs_p.extend(run_p.extra_records)
np.savez_compressed(os.path.join(out, "pusher_rollout.npz"), records=np.array(recs_p, np.float32))
print("pusher rollout golden: %d records, counters %s" % (len(recs_p), run_p.counters))
print("golden fixtures written to", out, [f for f in os.listdir(out)])
