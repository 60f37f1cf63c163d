"""Parity at the benchmarked scale and horizon:

  * the 4096-env SawyerPushObstacle-v0 bench configuration (push preset incl. reuse_data), a seeded 64-env subset followed
    through a full 250-step episode against the scalar restatement: state after every env.step, every record;
  * >= 10^6 state-validity queries against the f32 oracle (bit exact) and the f64 oracle (flips counted);
  * physics invariants that do not involve the oracle or its shared front end (mjcf.py / dynmodel.py): analytic free
    fall, torque-free spin, static contact force = weight, Coulomb deceleration.
"""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from helpers import planner_setup, random_qpos

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BENCH_SEED_ENV, BENCH_SEED_POLICY, BENCH_SEED_CFG = 1234, 1241, 1234   # bench.py: venv seed 1234, policy seed + 7, MoPAConfig.seed


TASKS = {   # task -> (env id, envs per GPU of the bench configuration, omega of scripts/3d/<task>/mopa.sh, action dims, subset size)
    "push": ("SawyerPushObstacle-v0", 4096, 0.7, 7, 64),
    "assembly": ("SawyerAssemblyObstacle-v0", 16384, 0.7, 7, 32),
    "lift": ("SawyerLiftObstacle-v0", 1024, 0.5, 8, 32),
}


def _task_cls(task):
    from mopa_rl_b200 import envs

    return {"push": envs.VecSawyerPushObstacle, "assembly": envs.VecSawyerAssemblyObstacle, "lift": envs.VecSawyerLiftObstacle}[task]


def _scalar_episode(args):
    """Worker: the scalar loop of one environment until its first episode ends; logs the state after every env.step."""
    gid, horizon = args[:2]
    task = args[2] if len(args) > 2 else "push"
    sys.path.insert(0, ROOT)
    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import MoPAConfig, env_planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    env_id, _, omega, adim, _ = TASKS[task]
    model = load_model(env_id)
    ignored, passive, _ = env_planner_inputs(_task_cls(task), model)
    cfg = MoPAConfig(max_iter=1000, reuse_data=True, max_reuse_data=15, seed=BENCH_SEED_CFG, omega=omega)

    def policy(g, k):
        u = rng.uniform01(BENCH_SEED_POLICY, np.uint64(g), np.uint64(k), np.arange(adim, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    r = ScalarMoPARunner(model, DynModel(model), cfg, ignored, passive, gid, BENCH_SEED_ENV, policy, max_episode_steps=horizon, task=task)
    states, env = {}, r.env
    step0, null0 = env.step, env.null_step

    def step(a, is_planner=False):
        out = step0(a, is_planner)
        states[env.ep_len] = (env.qpos.copy(), env.qvel.copy())
        return out

    def null_step():
        out = null0()
        states[env.ep_len] = (env.qpos.copy(), env.qvel.copy())
        return out

    env.step, env.null_step = step, null_step
    records = []
    while True:
        rec = r.macro_step()
        records.append(rec)
        records.extend(r.extra_records)
        if rec[49] == 1.0:
            break
    n = max(states)
    return gid, np.stack([states[k][0] for k in range(1, n + 1)]), np.stack([states[k][1] for k in range(1, n + 1)]), np.stack(records)


@pytest.mark.parametrize("task", ["push", "assembly", "lift"])
def test_bench_config_subset_matches_scalar_loop_over_a_full_episode(oracle_built, task):
    """push: BASELINE configs[1] (4096 envs); assembly: configs[3] (16384 envs, max_iter 1000); lift: the per-GPU share of
    configs[2] (1024 envs, omega 0.5)."""
    import torch

    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner

    _, n, omega, adim, nsub = TASKS[task]
    horizon, no = 250, _task_cls(task).OBS_DIM
    cfg = MoPAConfig(max_iter=1000, reuse_data=True, max_reuse_data=15, seed=BENCH_SEED_CFG, omega=omega)
    venv = _task_cls(task)(n, seed=BENCH_SEED_ENV, max_episode_steps=horizon)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, BENCH_SEED_POLICY, action_dim=adim),
                                     transition_capacity=max(1 << 20, n * 512))   # every record of the run must still be in the ring at the end
    subset = np.sort(np.random.default_rng(2026).choice(n, nsub, replace=False))
    sub_t = torch.as_tensor(subset, device=venv.dev)
    with mp.get_context("fork").Pool(min(nsub, os.cpu_count() or 1)) as pool:
        job = pool.map_async(_scalar_episode, [(int(g), horizon, task) for g in subset])
        # device side: tick until every subset env has finished its first episode; log (episode, ep_len, qpos, qvel) per tick
        logs = {int(g): {} for g in subset}
        episode = np.zeros(nsub, np.int64)
        prev_len = np.zeros(nsub, np.int64)
        for t in range(horizon * 2):
            runner.tick()
            ep_len = venv.ep_len[sub_t].cpu().numpy()
            q, v, work = venv.qpos[sub_t].cpu().numpy(), venv.qvel[sub_t].cpu().numpy(), venv.work[sub_t].cpu().numpy()
            for i, g in enumerate(subset):
                if ep_len[i] < prev_len[i]:
                    episode[i] += 1
                prev_len[i] = ep_len[i]
                if episode[i] == 0 and ep_len[i] > 0:
                    logs[int(g)][int(ep_len[i])] = (q[i], v[i], int(work[i]))
            if (episode >= 1).all():
                break
        assert (episode >= 1).all(), "some environments did not finish an episode in %d ticks" % (horizon * 2)
        runner.drain()
        torch.cuda.synchronize()
        c = runner.counters
        rec = runner.transitions[:c["transitions"]].cpu().numpy()
        scalar = job.get(timeout=1200)
    # Contact dynamics are chaotic: an env.step in which the arm is wedged against something (the Newton solver needs >= 2 steps in
    # every substep instead of 1) amplifies rounding-level differences (1e-15) by ten orders of magnitude within that one step - in
    # any two builds of the same physics, this kernel and its oracle included.  The 1e-5 bound is therefore asserted on every
    # env.step up to (not including) an environment's first such step; what happens afterwards is reported and bounded loosely.
    HEAVY = 2 * 75
    err_at = {1: 0.0, 75: 0.0, 250: 0.0}
    worst_q = worst_v = worst_obs = late_q = late_v = 0.0
    n_rec = n_steps = n_late = clean_envs = 0
    for gid, sq, sv, srec in scalar:
        log = logs[gid]
        # the last env.step of an episode is overwritten by the reset before the tick ends: compare what was logged
        ks = sorted(k for k in log if k <= len(sq))
        assert len(ks) >= len(sq) - 1, (gid, len(ks), len(sq))
        first_heavy = min([k for k in ks if log[k][2] >= HEAVY] + [10 ** 9])
        eq_all = max(np.abs(log[k][0] - sq[k - 1]).max() for k in ks)
        ev_all = max(np.abs(log[k][1] - sv[k - 1]).max() for k in ks)
        clean_envs += eq_all < 1e-5 and ev_all < 1e-5
        for k in ks:
            eq, ev = np.abs(log[k][0] - sq[k - 1]).max(), np.abs(log[k][1] - sv[k - 1]).max()
            if k < first_heavy:
                worst_q, worst_v = max(worst_q, eq), max(worst_v, ev)
                n_steps += 1
                if k in err_at:
                    err_at[k] = max(err_at[k], eq)
            else:
                late_q, late_v = max(late_q, eq), max(late_v, ev)
                n_late += 1
        mine = rec[rec[:, 51] == gid][:len(srec)]
        assert len(mine) == len(srec), (gid, len(mine), len(srec))
        # actions of relabelled records are displacements between visited states: they inherit the state error, so the
        # rounding-level bound only holds for environments without a heavy-contact step (see above)
        atol_ac = 1e-6 if first_heavy > len(sq) else 1e-3
        for k, (r, o) in enumerate(zip(mine, srec)):   # the macro-action structure agrees for every environment
            assert np.allclose(r[40:40 + adim], o[40:40 + adim], atol=atol_ac), (gid, k)
            assert r[49] == o[49] and r[50] == o[50], (gid, k, r[48:51], o[48:51])
            if first_heavy > len(sq):
                assert abs(r[48] - o[48]) < 1e-4, (gid, k, r[48], o[48])
                worst_obs = max(worst_obs, np.abs(r[0:no] - o[0:no]).max(), np.abs(r[52:52 + no] - o[52:52 + no]).max())
        n_rec += len(srec)
    print(task + " bench-config parity: %d envs x %d env.steps, %d records; %d env.steps before any heavy-contact step: max |dqpos| %.3e, max |dqvel| %.3e "
          "(after 1 / 75 / 250 env.steps: %.3e / %.3e / %.3e), max |dobs| %.3e; %d env.steps at / after a heavy-contact step: max |dqpos| %.3e, "
          "max |dqvel| %.3e; %d of %d environments within 1e-5 over the whole episode; counters %s"
          % (nsub, horizon, n_rec, n_steps, worst_q, worst_v, err_at[1], err_at[75], err_at[250], worst_obs, n_late, late_q, late_v, clean_envs, nsub,
             {k: c[k] for k in ("mp", "rl", "interpolation", "mp_fail", "reused", "fb_simple", "fb_main", "unstable")}))
    assert worst_q < 1e-5 and worst_v < 1e-5, (worst_q, worst_v)
    assert worst_obs < 1e-4
    assert clean_envs >= 0.9 * nsub and late_q < 5e-2, (clean_envs, late_q, late_v)


def test_single_substep_matches_oracle(push_model, oracle_built):
    """One mj_step (frame_dt = timestep): the '1 substep' point of BASELINE.md section 3, 256 envs."""
    import torch

    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerPushObstacle, push_reset_state
    from oracle.env_oracle import PushEnvOracle

    n = 256
    venv = VecSawyerPushObstacle(n, seed=5, frame_dt=0.002)
    venv.reset()
    dm = DynModel(push_model)
    q0, v0 = push_reset_state(push_model, 5, np.arange(n), np.zeros(n, dtype=np.int64))
    act = np.random.default_rng(1).uniform(-1, 1, (n, 8)).astype(np.float32)
    venv.step(torch.as_tensor(act, device="cuda"))
    torch.cuda.synchronize()
    gq, gv = venv.qpos.cpu().numpy(), venv.qvel.cpu().numpy()
    worst = 0.0
    for i in range(n):
        e = PushEnvOracle(push_model, dm, frame_dt=0.002)
        e.reset_to(q0[i], v0[i])
        e.step(act[i].astype(np.float64), False)
        worst = max(worst, np.abs(gq[i] - e.qpos).max(), np.abs(gv[i] - e.qvel).max())
    print("one substep, %d envs: max |dq|, |dv| = %.3e" % (n, worst))
    assert worst < 1e-9


def test_million_states_bit_exact_and_f64_flips(push_model, oracle_built):
    """>= 10^6 validity queries: words bit-identical to the f32 oracle; against the f64 oracle (the
    reference computes in double) the number of flipped booleans is counted and bounded."""
    import threading

    from mopa_rl_b200.capi import NativePlanner

    ignored, passive, ref = planner_setup(push_model)
    native = NativePlanner(push_model, passive, ignored, -0.002, 0.1, seed=1234)
    n = 1_048_576
    q = random_qpos(push_model, n, 4242, ref)
    _, w = native.is_valid_host(q, flags=1, return_words=True)
    threads = min(os.cpu_count() or 1, 32)
    parts = np.array_split(np.arange(n), threads)
    out32, out64 = [None] * threads, [None] * threads

    def work(i):
        out32[i] = oracle_built.OracleScene(push_model, ignored, -0.002, "f32").is_valid(q[parts[i]])
        out64[i] = oracle_built.OracleScene(push_model, ignored, -0.002, "f64").is_valid(q[parts[i]])

    th = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    o32, o64 = np.concatenate(out32), np.concatenate(out64)
    assert np.array_equal(w, o32), "%d of %d words differ from the f32 oracle" % ((w != o32).sum(), n)
    flips = int(((w & 1) != (o64 & 1)).sum())
    print("validity: %d states, valid fraction %.4f, f32 words bit-exact, booleans flipped against the f64 oracle: %d" % (n, (w & 1).mean(), flips))
    assert flips <= 8, flips   # states within fp32 rounding of the contact threshold


# ----------------------------------------------------------------------------- oracle-independent physics invariants
def _cube_env(n, push_model, **kw):
    import torch

    from mopa_rl_b200.envs import VecSawyerPushObstacle

    venv = VecSawyerPushObstacle(n, seed=1, max_episode_steps=10 ** 6, **kw)
    venv.reset()
    a = push_model.get_joint_qpos_addr("cube")[0]
    va = push_model.get_joint_qvel_addr("cube")[0]
    return venv, a, va, torch


# constants read by hand from env/assets/xml/sawyer_push_obstacle.xml:46-48 (not through mjcf.py / dynmodel.py):
#   <geom name="cube" type="box" size="0.03 0.03 0.03" density="300" friction="0.95 ..."/>  <joint name="cube" type="free" damping="0.0005"/>
CUBE_MASS = 300.0 * 0.06 ** 3
CUBE_INERTIA = CUBE_MASS * 0.06 ** 2 / 6.0
CUBE_DAMPING = 0.0005


def test_free_fall_is_the_semi_implicit_euler_recurrence(push_model):
    """A cube released in mid air.  mj_Euler with implicit joint damping: (m + h d) v' = m v - h m g, z' = z + h v' - the closed
    recurrence is evaluated here from the XML's numbers; a torque-free spinning cube (isotropic inertia: no precession)
    keeps its axis, loses spin at I / (I + h d) per step, and its quaternion stays normalised."""
    venv, a, va, torch = _cube_env(4, push_model)
    q, v = venv.qpos.clone(), venv.qvel.clone()
    z0 = 1.6
    q[:, a:a + 3] = torch.tensor([0.6, 0.3, z0], dtype=torch.float64, device=venv.dev)
    w0 = np.array([[0, 0, 0], [3.0, 0, 0], [1.0, -2.0, 0.5], [0, 0, 7.0]])
    v[:, va + 3:va + 6] = torch.as_tensor(w0, device=venv.dev)
    venv.set_state(np.arange(4), q.cpu().numpy(), v.cpu().numpy())
    h, g, nsub = 0.002, 9.81, 75
    z, vz, spin = z0, 0.0, 1.0
    for step in range(1, 3):
        venv.step(torch.zeros(4, 8, device=venv.dev))
        torch.cuda.synchronize()
        for _ in range(nsub):
            vz = (CUBE_MASS * vz - h * CUBE_MASS * g) / (CUBE_MASS + h * CUBE_DAMPING)
            z += h * vz
            spin *= CUBE_INERTIA / (CUBE_INERTIA + h * CUBE_DAMPING)
        gz, gvz = venv.qpos[:, a + 2].cpu().numpy(), venv.qvel[:, va + 2].cpu().numpy()
        assert np.abs(gvz - vz).max() < 1e-10, (gvz, vz)
        assert np.abs(gz - z).max() < 1e-10, (gz, z)
        quat = venv.qpos[:, a + 3:a + 7].cpu().numpy()
        assert np.abs(np.linalg.norm(quat, axis=1) - 1).max() < 1e-12
        w = venv.qvel[:, va + 3:va + 6].cpu().numpy()
        assert np.abs(w - spin * w0).max() < 1e-9, (w, spin * w0)
        assert np.abs(venv.qvel[:, va:va + 2].cpu().numpy()).max() < 1e-12
    assert int(venv.ncon.max()) == 0
    print("free fall: z, vz after 150 mj_steps match the closed recurrence to %.1e / %.1e" % (np.abs(gz - z).max(), np.abs(gvz - vz).max()))


def test_resting_cube_is_carried_by_its_weight(push_model):
    """Static equilibrium on the bin floor: contact complementarity (no penetration velocity, contacts active) and the
    sum of the normal forces = m g; friction idle."""
    venv, a, va, torch = _cube_env(8, push_model)
    for _ in range(6):                                   # let the soft contact settle (0.9 s)
        venv.step(torch.zeros(8, 8, device=venv.dev))
    torch.cuda.synchronize()
    weight = CUBE_MASS * 9.81
    cf = venv.cforce.cpu().numpy()
    assert int(venv.ncon.min()) >= 1
    assert np.abs(venv.qvel[:, va:va + 6].cpu().numpy()).max() < 1e-6
    assert np.abs(cf - weight).max() < 1e-3 * weight, (cf, weight)


def test_sliding_cube_decelerates_at_mu_g(push_model):
    """Coulomb friction: a cube sliding on the bin floor loses speed at about mu g while it slides (mu = max of the two
    geoms' friction coefficients, as MuJoCo combines them)."""
    venv, a, va, torch = _cube_env(2, push_model, frame_dt=0.05)   # 25 mj_steps per env.step
    for _ in range(20):
        venv.step(torch.zeros(2, 8, device=venv.dev))
    v = venv.qvel.clone()
    v0 = 1.0
    v[0, va] = v0                                        # env 0 slides along x, env 1 stays
    venv.set_state(np.arange(2), venv.qpos.cpu().numpy(), v.cpu().numpy())
    venv.step(torch.zeros(2, 8, device=venv.dev))
    torch.cuda.synchronize()
    mu = 1.0   # max(cube 0.95, bin floor 1.0 - MuJoCo's default friction, the bin geoms set none)
    vx = float(venv.qvel[0, va])
    decel = (v0 - vx) / 0.05
    print("sliding cube: v after 0.05 s = %.4f, deceleration %.2f m/s^2 (Coulomb: mu g = %.2f)" % (vx, decel, mu * 9.81))
    # mu = 1 on a cube is the tipping limit (friction torque = restoring torque): the load shifts to the leading edge and the
    # normal force overshoots m g while the cube pitches, so the deceleration is only required to be mu g within 25 %
    assert 0.75 * mu * 9.81 < decel < 1.25 * mu * 9.81, (vx, decel)
    assert abs(float(venv.qvel[1, va])) < 1e-6
