"""Conformance of the vectorised MoPA runner against the scalar restatement of
MoPARolloutRunner.run (oracle/rollout_oracle.py): same policy draws, same planner keys -> the
per-env sequences of SMDP transition records must agree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_vectorised_runner_matches_scalar_reference_loop(push_model, oracle_built):
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    n, ticks, seed = 12, 60, 4321
    cfg = MoPAConfig(max_iter=150, seed=99)
    venv = VecSawyerPushObstacle(n, seed=seed, max_episode_steps=25, env_id_offset=100)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 7))
    for _ in range(ticks):
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    rec = runner.transitions[:runner.n_transitions].cpu().numpy()
    assert n * ticks * 0.8 <= runner.env_steps and len(rec) > n   # envs waiting for an RRT plan skip ticks

    def policy(gid, k):
        u = rng.uniform01(7, np.uint64(gid), np.uint64(k), np.arange(7, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ignored, passive, _ = planner_inputs(push_model)
    dm = DynModel(push_model)
    kinds = set()
    worst = 0.0
    for e in range(n):
        gid = 100 + e
        mine = rec[rec[:, 51] == gid]
        ref = ScalarMoPARunner(push_model, dm, cfg, ignored, passive, gid, seed, policy, max_episode_steps=25)
        for k, r in enumerate(mine):
            o = ref.macro_step()
            assert np.array_equal(r[40:47], o[40:47]), (e, k)                    # same action
            assert r[49] == o[49] and r[50] == o[50], (e, k, r[48:51], o[48:51])  # done flag, intra_steps
            assert abs(r[48] - o[48]) < 1e-5, (e, k)                              # (discounted) reward
            d = max(np.abs(r[0:40] - o[0:40]).max(), np.abs(r[52:92] - o[52:92]).max())
            worst = max(worst, d)
            assert d < 1e-4, (e, k, d)
            kinds.add("plan" if r[50] > 0 else "single")
    assert kinds == {"plan", "single"}
    print("vectorised vs scalar runner: %d transitions, worst |obs diff| %.2e, counters %s" % (len(rec), worst, runner.counters))


def test_native_runner_assembly_matches_scalar_reference_loop(oracle_built):
    """Same conformance check on SawyerAssemblyObstacle-v0 (BASELINE configs[3]: planner-heavy scene, 39 collidable
    geoms, ignored pairs = furniture parts x table)."""
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerAssemblyObstacle
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, env_planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    model = load_model("SawyerAssemblyObstacle-v0")
    n, ticks, seed = 8, 40, 977
    cfg = MoPAConfig(max_iter=150, seed=5)
    venv = VecSawyerAssemblyObstacle(n, seed=seed, max_episode_steps=20, env_id_offset=40)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 11))
    for _ in range(ticks):
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    rec = runner.transitions[:runner.n_transitions].cpu().numpy()
    assert len(rec) > n

    def policy(gid, k):
        u = rng.uniform01(11, np.uint64(gid), np.uint64(k), np.arange(7, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ignored, passive, _ = env_planner_inputs(VecSawyerAssemblyObstacle, model)
    dm = DynModel(model)
    worst = 0.0
    for e in range(n):
        gid = 40 + e
        mine = rec[rec[:, 51] == gid]
        ref = ScalarMoPARunner(model, dm, cfg, ignored, passive, gid, seed, policy, max_episode_steps=20, task="assembly")
        for k, r in enumerate(mine):
            o = ref.macro_step()
            assert np.array_equal(r[40:47], o[40:47]), (e, k)
            assert r[49] == o[49] and r[50] == o[50], (e, k, r[48:51], o[48:51])
            assert abs(r[48] - o[48]) < 1e-5, (e, k)
            d = max(np.abs(r[0:40] - o[0:40]).max(), np.abs(r[52:92] - o[52:92]).max())
            worst = max(worst, d)
            assert d < 1e-4, (e, k, d)
    print("assembly: native vs scalar runner: %d transitions, worst |obs diff| %.2e, counters %s" % (len(rec), worst, runner.counters))


def test_native_runner_reuse_data_matches_scalar_reference_loop(push_model, oracle_built):
    """reuse_data (scripts/3d/push/mopa.sh: reuse_data=True, max_reuse_data=15; rl/mopa_rollouts.py:223-302): every
    executed plan of more than 3 steps also yields relabelled (start, goal) transitions.  Main and relabelled records
    of every environment must match the scalar restatement, in order."""
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    n, ticks, seed = 10, 50, 2024
    cfg = MoPAConfig(max_iter=150, seed=17, reuse_data=True, max_reuse_data=15)
    venv = VecSawyerPushObstacle(n, seed=seed, max_episode_steps=30, env_id_offset=7)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 3))
    for _ in range(ticks):
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    c = runner.counters
    rec = runner.transitions[:c["transitions"]].cpu().numpy()
    assert c["reused"] > n

    def policy(gid, k):
        u = rng.uniform01(3, np.uint64(gid), np.uint64(k), np.arange(7, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ignored, passive, _ = planner_inputs(push_model)
    dm = DynModel(push_model)
    n_extra, worst = 0, 0.0
    for e in range(n):
        gid = 7 + e
        mine = list(rec[rec[:, 51] == gid])
        ref = ScalarMoPARunner(push_model, dm, cfg, ignored, passive, gid, seed, policy, max_episode_steps=30)
        k = 0
        while k < len(mine):
            expect = [ref.macro_step()] + list(ref.extra_records)
            n_extra += len(expect) - 1
            for o in expect:
                if k >= len(mine):
                    break   # the collection stopped between a main record and its relabelled ones
                r = mine[k]
                assert np.allclose(r[40:47], o[40:47], atol=1e-6), (e, k, r[40:47], o[40:47])
                assert r[49] == o[49] and r[50] == o[50], (e, k, r[48:51], o[48:51])
                assert abs(r[48] - o[48]) < 1e-5, (e, k)
                d = max(np.abs(r[0:40] - o[0:40]).max(), np.abs(r[52:92] - o[52:92]).max())
                worst = max(worst, d)
                assert d < 1e-4, (e, k, d)
                k += 1
    assert n_extra > n
    print("reuse_data: %d records (%d relabelled), worst |obs diff| %.2e" % (len(rec), c["reused"], worst))


def test_run_episodes_matches_scalar_episodes(push_model, oracle_built):
    """Evaluation path (run_episode, rl/mopa_rollouts.py:401-681): per-episode len / rew / success / contact_force of every
    environment against the scalar restatement run for the same number of episodes."""
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, planner_inputs, run_episodes
    from oracle.rollout_oracle import ScalarMoPARunner

    n, horizon, episodes, seed = 8, 12, 2, 11
    cfg = MoPAConfig(max_iter=100, seed=2)
    venv = VecSawyerPushObstacle(n, seed=seed, max_episode_steps=horizon)
    info = run_episodes(venv, cfg, policy=CounterPolicy(torch, venv.dev, 5), episodes_per_env=episodes)
    assert info["episodes"] >= episodes * n
    assert info["mp"] + info["rl"] + info["interpolation"] + info["mp_fail"] > 0

    def policy(gid, k):
        u = rng.uniform01(5, np.uint64(gid), np.uint64(k), np.arange(7, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ignored, passive, _ = planner_inputs(push_model)
    dm = DynModel(push_model)
    per = info["per_env"]
    for e in range(n):
        ref = ScalarMoPARunner(push_model, dm, cfg, ignored, passive, e, seed, policy, max_episode_steps=horizon)
        done_eps, tot_len, tot_rew, tot_succ = 0, 0, 0.0, 0.0
        target = int(per[e, 0])                          # the device may have finished more than `episodes` before the last check
        while done_eps < target:
            rec = ref.macro_step()
            if rec[49] == 1.0:
                done_eps += 1
                tot_len += ref.env.ep_len
                tot_rew += ref.env.ep_rew
                tot_succ += float(ref.env.success)
        assert per[e, 1] == tot_len, (e, per[e], tot_len)
        assert abs(per[e, 2] - tot_rew) < 1e-6 and per[e, 3] == tot_succ, (e, per[e], tot_rew, tot_succ)
        assert abs(per[e, 4] - ref.contact_force_sum) <= 1e-6 * max(1.0, ref.contact_force_sum), (e, per[e, 4], ref.contact_force_sum)
    # sanity of the aggregate: the cube rests on the bin floor in (almost) every step, contact force per step ~ its weight
    weight = push_model.body_mass[push_model.body_name2id("cube")] * 9.81
    assert 0.5 * weight * horizon < info["contact_force"] < 3.0 * weight * horizon, (info["contact_force"], weight)
    print("run_episodes: %d episodes, mean len %.1f rew %.4f contact force %.4f - per-env sums equal the scalar loop's" % (info["episodes"], info["len"], info["rew"], info["contact_force"]))


@pytest.mark.parametrize("ac_space_type", ["piecewise", "normal"])
def test_native_runner_discrete_action_matches_scalar_reference_loop(push_model, oracle_built, ac_space_type):
    """config.discrete_action (scripts/3d/push/mopa_discrete.sh: omega = 0, reuse_data): the policy's ac_type picks
    motion planner / direct execution (rl/mopa_rollouts.py:86-88, 104-111), direct actions are executed unscaled
    (:347-352), relabelled records inherit ac_type (:266-267).  Record slot 47 carries ac_type.  ac_space_type "normal"
    (the lift / assembly / 2-D discrete presets): displacement = a * action_range, relabelled action = d / action_range."""
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    n, ticks, seed = 10, 50, 606
    cfg = MoPAConfig(max_iter=150, seed=23, omega=0.0, discrete_action=True, reuse_data=True, max_reuse_data=15, ac_space_type=ac_space_type)
    venv = VecSawyerPushObstacle(n, seed=seed, max_episode_steps=30, env_id_offset=300)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 13, discrete=True))
    for _ in range(ticks):
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    c = runner.counters
    rec = runner.transitions[:c["transitions"]].cpu().numpy()
    assert c["rl"] > n and c["mp"] + c["interpolation"] > n

    def policy(gid, k):
        u = rng.uniform01(13, np.uint64(gid), np.uint64(k), np.arange(8, dtype=np.uint64))
        return (2.0 * u[:7] - 1.0).astype(np.float32), bool(u[7] < 0.5)

    ignored, passive, _ = planner_inputs(push_model)
    dm = DynModel(push_model)
    types, worst = set(), 0.0
    for e in range(n):
        gid = 300 + e
        mine = list(rec[rec[:, 51] == gid])
        ref = ScalarMoPARunner(push_model, dm, cfg, ignored, passive, gid, seed, policy, max_episode_steps=30)
        k = 0
        while k < len(mine):
            for o in [ref.macro_step()] + list(ref.extra_records):
                if k >= len(mine):
                    break
                r = mine[k]
                assert np.allclose(r[40:47], o[40:47], atol=1e-6) and r[47] == o[47], (e, k, r[40:48], o[40:48])
                assert r[49] == o[49] and r[50] == o[50], (e, k, r[48:51], o[48:51])
                assert abs(r[48] - o[48]) < 1e-5, (e, k)
                d = max(np.abs(r[0:40] - o[0:40]).max(), np.abs(r[52:92] - o[52:92]).max())
                worst = max(worst, d)
                assert d < 1e-4, (e, k, d)
                types.add(float(r[47]))
                k += 1
    assert types == {0.0, 1.0}
    print("discrete_action (%s): %d records, worst |obs diff| %.2e, counters %s" % (ac_space_type, len(rec), worst, c))


def test_native_runner_lift_matches_scalar_reference_loop(oracle_built):
    """SawyerLiftObstacle-v0 (BASELINE configs[2] scene): mesh collider in the planner, 8-D actions whose gripper entry
    is executed with direct actions and with the last waypoint of a plan (rl/mopa_rollouts.py:170-175), form_action's
    gripper entry in between (env/sawyer/sawyer.py:290-296), reuse_data on."""
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerLiftObstacle
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, env_planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    model = load_model("SawyerLiftObstacle-v0")
    n, ticks, seed = 8, 40, 515
    cfg = MoPAConfig(max_iter=150, seed=31, reuse_data=True, max_reuse_data=15)
    venv = VecSawyerLiftObstacle(n, seed=seed, max_episode_steps=20, env_id_offset=60)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 19, action_dim=8))
    for _ in range(ticks):
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    c = runner.counters
    rec = runner.transitions[:c["transitions"]].cpu().numpy()
    assert len(rec) > n

    def policy(gid, k):
        u = rng.uniform01(19, np.uint64(gid), np.uint64(k), np.arange(8, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ignored, passive, _ = env_planner_inputs(VecSawyerLiftObstacle, model)
    dm = DynModel(model)
    worst, n_plan = 0.0, 0
    for e in range(n):
        gid = 60 + e
        mine = list(rec[rec[:, 51] == gid])
        ref = ScalarMoPARunner(model, dm, cfg, ignored, passive, gid, seed, policy, max_episode_steps=20, task="lift")
        k = 0
        while k < len(mine):
            for o in [ref.macro_step()] + list(ref.extra_records):
                if k >= len(mine):
                    break
                r = mine[k]
                assert np.allclose(r[40:48], o[40:48], atol=1e-6), (e, k, r[40:48], o[40:48])
                assert r[49] == o[49] and r[50] == o[50], (e, k, r[48:51], o[48:51])
                assert abs(r[48] - o[48]) < 1e-5, (e, k)
                d = max(np.abs(r[0:35] - o[0:35]).max(), np.abs(r[52:87] - o[52:87]).max())
                worst = max(worst, d)
                assert d < 1e-4, (e, k, d)
                n_plan += r[50] > 0
                k += 1
    assert n_plan > n
    print("lift: native vs scalar runner: %d records, worst |obs diff| %.2e, counters %s" % (len(rec), worst, c))


@pytest.mark.parametrize("omega", [0.05, 3.0], ids=["preset-omega-holds", "large-omega-direct"])
def test_ik_target_rollout_matches_scalar_reference_loop(oracle_built, omega):
    """config.use_ik_target (scripts/3d/lift/mopa_ik.sh: omega 0.05, action_range 0.2, ik_target grip_site): the policy acts in
    Cartesian space, (default[3], quat[4], gripper); every macro action runs _cart2dispalcement (rl/mopa_rollouts.py:683-728) -
    mat2quat of the site frame with the reference's [3, 0, 1, 1] index list, qpos_from_site_pose(max_steps 100, tol 1e-2) -
    and the joint displacement decides: above omega the "planner" branch, whose target is the current state in the reference
    (a two-step hold, gripper entry on the second step), else env.step(displacement / omega + gripper).  With the preset's
    omega every action of a random policy ends in the hold (the [3, 0, 1, 1] orientation target is never the current one, the
    wrist always has to turn by more than 0.05 rad); a large omega sends most actions down the direct branch.  The scalar
    side takes the reference's eigendecomposition route through mat2quat, the kernel the closed form: they agree to float32
    resolution of the matrix."""
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerLiftObstacle
    from mopa_rl_b200.model import load_model
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, env_planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    model = load_model("SawyerLiftObstacle-v0")
    n, ticks, seed, off = 8, 36, 77, 20
    cfg = MoPAConfig(omega=omega, action_range=0.2, max_iter=150, seed=9, reuse_data=True, max_reuse_data=15, use_ik_target=True)
    venv = VecSawyerLiftObstacle(n, seed=seed, max_episode_steps=16, env_id_offset=off)
    base = CounterPolicy(torch, venv.dev, 23, action_dim=8)

    def device_policy(obs, gid, mi):
        a = base(obs, gid, mi).clone()
        small = (mi % 2 == 0)
        a[small, :3] = a[small, :3] * 0.04
        return a

    runner = NativeMoPARolloutRunner(venv, cfg, policy=device_policy)
    for _ in range(ticks):
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    c = runner.counters
    rec = runner.transitions[:c["transitions"]].cpu().numpy()
    assert c["mp"] == 0 and c["reused"] == 0 and c["mp_fail"] == 0, c   # IK "plans" are two-step holds: nothing to relabel, no RRT
    assert (c["interpolation"] > n and c["rl"] == 0) if omega < 1 else (c["rl"] > n), c   # (some wrist turns exceed even 3 rad)

    def policy(gid, k):
        u = rng.uniform01(23, np.uint64(gid), np.uint64(k), np.arange(8, dtype=np.uint64))
        a = (2.0 * u - 1.0).astype(np.float32)
        if k % 2 == 0:
            a[:3] = a[:3] * np.float32(0.04)
        return a

    ignored, passive, _ = env_planner_inputs(VecSawyerLiftObstacle, model)
    dm = DynModel(model)
    worst, n_direct, n_hold = 0.0, 0, 0
    for e in range(n):
        gid = off + e
        mine = rec[rec[:, 51] == gid]
        ref = ScalarMoPARunner(model, dm, cfg, ignored, passive, gid, seed, policy, max_episode_steps=16, task="lift")
        for k, r in enumerate(mine):
            o = ref.macro_step()
            assert np.array_equal(r[40:48], o[40:48]), (e, k, r[40:48], o[40:48])          # the record keeps the Cartesian action
            assert r[49] == o[49] and r[50] == o[50], (e, k, r[48:51], o[48:51])
            assert abs(r[48] - o[48]) < 1e-4, (e, k, r[48], o[48])
            d = max(np.abs(r[0:35] - o[0:35]).max(), np.abs(r[52:87] - o[52:87]).max())
            worst = max(worst, d)
            assert d < 2e-4, (e, k, d)
            n_direct += r[50] == 0
            n_hold += r[50] == 1
    assert (n_hold > n and n_direct == 0) if omega < 1 else (n_direct > n), (n_direct, n_hold)
    print("lift IK: native vs scalar runner: %d records (%d direct, %d holds), worst |obs diff| %.2e, counters %s" % (len(rec), n_direct, n_hold, worst, c))


@pytest.mark.parametrize("simple_max_iter", [3, 0], ids=["simple-planner", "main-planner-retry"])
def test_blocked_hops_go_through_the_simple_and_the_main_planner(push_model, oracle_built, simple_max_iter):
    """SACAgent.simple_interpolate(use_planner=True) (rl/sac_agent.py:300-311): a densification hop whose interior is blocked is
    re-planned with the simple planner (range 0.05), then with the main planner, else only its end point is kept.  Such hops
    are rare (~0.3 % of the RRT plans), so the test hook `debug_block_mod` forces every third hop down that path, in the
    kernels and in the scalar restatement alike; the executed trajectories (records) must still agree."""
    import torch

    from mopa_rl_b200 import rng
    from mopa_rl_b200.dynmodel import DynModel
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.rollout import CounterPolicy, MoPAConfig, NativeMoPARolloutRunner, planner_inputs
    from oracle.rollout_oracle import ScalarMoPARunner

    n, ticks, seed = 256, 60, 811
    # simple_max_iter = 3: the simple planner (range 0.05) solves the hops; 0: it gives up at once and the main planner is retried
    cfg = MoPAConfig(max_iter=150, seed=41, debug_block_mod=3, simple_max_iter=simple_max_iter)
    venv = VecSawyerPushObstacle(n, seed=seed, max_episode_steps=40, env_id_offset=500)
    runner = NativeMoPARolloutRunner(venv, cfg, policy=CounterPolicy(torch, venv.dev, 29))
    for _ in range(ticks):
        runner.tick()
    runner.drain()
    torch.cuda.synchronize()
    c = runner.counters
    rec = runner.transitions[:c["transitions"]].cpu().numpy()
    assert c["fb_simple" if simple_max_iter else "fb_main"] >= 5 and c["fb_main" if simple_max_iter else "fb_simple"] == 0, c

    def policy(gid, k):
        u = rng.uniform01(29, np.uint64(gid), np.uint64(k), np.arange(7, dtype=np.uint64))
        return (2.0 * u - 1.0).astype(np.float32)

    ignored, passive, _ = planner_inputs(push_model)
    dm = DynModel(push_model)
    # RRT plans are ~2 % of the macro actions: follow the environments that executed one (a straight-line plan has at most
    # 13 steps, intra_steps <= 12), plus a few others
    with_rrt = sorted(set(rec[rec[:, 50] > 12][:, 51].astype(int)))
    chosen = with_rrt[:20] + [500 + e for e in range(0, n, 64)]
    assert len(with_rrt) >= 3, c
    worst, tot = 0.0, dict(fb_simple=0, fb_main=0, densify_fallback=0, mp=0)
    for gid in chosen:
        mine = rec[rec[:, 51] == gid]
        ref = ScalarMoPARunner(push_model, dm, cfg, ignored, passive, gid, seed, policy, max_episode_steps=40)
        for k, r in enumerate(mine):
            o = ref.macro_step()
            assert np.array_equal(r[40:47], o[40:47]), (gid, k)
            assert r[49] == o[49] and r[50] == o[50], (gid, k, r[48:51], o[48:51])   # intra_steps = length of the executed trajectory
            assert abs(r[48] - o[48]) < 1e-5, (gid, k)
            d = max(np.abs(r[0:40] - o[0:40]).max(), np.abs(r[52:92] - o[52:92]).max())
            worst = max(worst, d)
            assert d < 1e-4, (gid, k, d)
        for key in tot:
            tot[key] += ref.counters[key]
    assert tot["mp"] >= 3 and tot["fb_simple"] + tot["fb_main"] > 0, tot
    print("fallback planners: device %s; scalar loops of %d envs %s; worst |obs diff| %.2e" % ({k: c[k] for k in tot}, len(chosen), tot, worst))


def test_limits_are_refused_not_clamped(push_model):
    """max_reuse_data beyond the relabelling kernel's capacity and action ranges beyond the straight-line planner's 16
    interpolation points raise (round 1 clamped silently)."""
    from mopa_rl_b200.envs import VecSawyerPushObstacle
    from mopa_rl_b200.rollout import MoPAConfig, NativeMoPARolloutRunner

    venv = VecSawyerPushObstacle(4, seed=1)
    with pytest.raises(NotImplementedError):
        NativeMoPARolloutRunner(venv, MoPAConfig(reuse_data=True, max_reuse_data=33))
    with pytest.raises(NotImplementedError):
        NativeMoPARolloutRunner(venv, MoPAConfig(action_range=1.0, ac_scale=0.05))
    r = NativeMoPARolloutRunner(venv, MoPAConfig(reuse_data=True, max_reuse_data=30))   # scripts/2d/mopa.sh
    r.tick()
    r.close()
