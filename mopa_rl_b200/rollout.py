"""Vectorised MoPA experience collection — the batched replacement of ``MoPARolloutRunner.run``
(rl/mopa_rollouts.py:22-399) together with the planner glue it calls in ``SACAgent``
(rl/sac_agent.py:145-318: is_planner_ac, convert2planner_displacement, clip_qpos,
simple_interpolate, plan).

The reference drives ONE env through a Python generator.  Here N envs live on the GPU and every
``tick()`` advances each of them by exactly one ``env.step``: an env either executes the next
waypoint of its current plan (``is_planner=True``), or — when it has no plan left — receives a
fresh policy action, which is executed directly (|a| <= omega) or turned into a plan by the same
sequence the reference uses: displacement map -> target clip -> invalid-target back-off ->
straight-line interpolation -> RRT-Connect -> densification.  All validity checks of a tick are
batched into a few ``mopa_is_valid_batch`` launches and all RRT problems into one
``mopa_plan_batch`` launch.  A finished macro action emits one SMDP transition record
(ob 40, ac 8, rew, done, intra_steps, env id, ob_next 40 = 92 floats), exactly the content of the
reference's ``Rollout`` entries (rl/rollouts.py:15-36).
"""
from __future__ import annotations

import numpy as np

from .capi import NativePlanner

TRANSITION_FLOATS = 92


class MoPAConfig:
    """Hyper-parameters of scripts/3d/push/mopa.sh + config/sawyer.py + config/motion_planner.py."""
    omega = 0.7
    action_range = 0.5
    ac_scale = 0.05            # SawyerEnv._ac_scale
    discount_factor = 0.99
    invalid_target_handling = True
    num_trials = 100
    step_size = 0.02
    joint_margin = 0.001
    interpolation = True
    contact_threshold = -0.002
    range = 0.1
    simple_planner_range = 0.05
    max_iter = 1000            # iteration cap standing in for --timelimit (2.0 s in the push preset)
    simple_max_iter = 25       # ... and for --simple_planner_timelimit (0.05 s, config/sawyer.py:92-96): max_iter * 0.05 / 2.0
    max_path = 384
    max_traj = 1280
    seed = 1234
    reuse_data = False         # scripts/3d/push/mopa.sh sets True: relabel (start, goal) pairs of every executed plan
    max_reuse_data = 15
    ac_space_type = "piecewise"  # "normal" (lift / assembly / 2d mopa_discrete.sh): displacement = a * action_range
    discrete_action = False    # scripts/3d/*/mopa_discrete.sh (with omega = 0): the policy's ac_type picks planner / direct
    debug_block_mod = 0        # test hook (mopa_rollout_config.debug_block_mod): force densification hops into the fallback planners
    use_ik_target = False      # scripts/3d/*/mopa_ik.sh: Cartesian actions (default[3], quat[4], gripper) through qpos_from_site_pose
    ik_target = "grip_site"
    ik_max_steps = 100         # rl/mopa_rollouts.py:706-707
    ik_tol = 1e-2

    def __init__(self, **kw):
        for k, v in kw.items():
            if not hasattr(type(self), k):
                raise AttributeError(k)
            setattr(self, k, v)


def planner_inputs(model, static_bodies=("table", "bin1"), manipulation_geoms=("cube",), manipulation_bodies=None):
    """ignored contact pairs / passive joints as rl/trainer.py:62-75 derives them: every geom of the manipulation
    bodies (env.manipulation_geom_ids) against every geom of the static bodies (env.static_geom_ids)."""
    body_of = lambda g: model.names["body"][model.geom_bodyid[g]]
    static_ids = [g for g in range(model.ngeom) if body_of(g) in static_bodies]
    if manipulation_bodies is not None:
        manip = [g for g in range(model.ngeom) if body_of(g) in manipulation_bodies]
    else:
        manip = [model.geom_name2id(name) for name in manipulation_geoms]
    ignored = []
    for mg in manip:
        ignored += [(min(mg, g), max(mg, g)) for g in static_ids]
    ref = [model.get_joint_qpos_addr("right_j%d" % i) for i in range(7)]
    passive = [i for i in range(model.nq) if i not in ref]
    return ignored, passive, ref


def env_planner_inputs(venv_cls, model):
    """planner_inputs for a vectorised env class (STATIC_BODIES / MANIPULATION_BODIES, or its own planner_inputs)."""
    if hasattr(venv_cls, "planner_inputs"):
        return venv_cls.planner_inputs(model)
    return planner_inputs(model, static_bodies=venv_cls.STATIC_BODIES, manipulation_bodies=venv_cls.MANIPULATION_BODIES)


class UniformPolicy:
    """random_exploration=True: ac_space.sample(), uniform in [-1, 1]^7 (rl/base_agent.py:15-22)."""

    def __init__(self, torch, device, seed, action_dim=7):
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(int(seed))
        self.torch, self.device, self.adim = torch, device, int(action_dim)

    def __call__(self, obs, env_ids=None, macro_index=None):
        return self.torch.rand(obs.shape[0], self.adim, generator=self.gen, device=self.device) * 2 - 1


class CounterPolicy:
    """Uniform [-1,1]^7 actions that are a pure function of (seed, env id, macro-action index):
    the stand-in for random_exploration whose draws do not depend on batch composition."""

    def __init__(self, torch, device, seed, discrete=False, action_dim=7):
        self.torch, self.device, self.seed, self.discrete, self.adim = torch, device, int(seed), discrete, int(action_dim)

    def __call__(self, obs, env_ids, macro_index):
        from . import rng

        e = env_ids.cpu().numpy().astype(np.uint64)[:, None]
        c = macro_index.cpu().numpy().astype(np.uint64)[:, None]
        u = rng.uniform01(self.seed, e, c, np.arange(self.adim, dtype=np.uint64)[None, :])
        ac = self.torch.as_tensor((2.0 * u - 1.0).astype(np.float32), device=self.device)
        if not self.discrete:
            return ac
        # discrete_action: ac_type ~ Discrete(2) from draw 7 of the same stream (ac_space.spaces["ac_type"], rl/trainer.py:90-91)
        t = rng.uniform01(self.seed, e[:, 0], c[:, 0], np.uint64(7)) < 0.5
        return ac, self.torch.as_tensor(t.astype(np.uint8), device=self.device)


# ------------------------------------------------------------------------------------ native runner
import ctypes as _C  # noqa: E402

COUNTER_NAMES = ("mp", "rl", "interpolation", "mp_fail", "approximate", "invalid", "densify_fallback", "episodes", "success",
                 "mp_path_len", "interpolation_path_len", "env_steps", "transitions", "rrt_dropped", "rrt_problems", "waiting", "reused", "unstable",
                 "fb_simple", "fb_main")


class _RolloutConfig(_C.Structure):
    _fields_ = [(k, _C.c_int32) for k in ("n_envs", "max_iter", "max_path", "max_traj", "rrt_capacity", "num_trials",
                                          "invalid_target_handling", "interpolation")] + \
               [(k, _C.c_double) for k in ("omega", "action_range", "ac_scale", "discount", "step_size", "joint_margin", "range")] + \
               [("seed_env", _C.c_uint64), ("env_id_offset", _C.c_int64), ("jnt_lo", _C.c_double * 7), ("jnt_hi", _C.c_double * 7),
                ("init_qpos", _C.c_double * 7), ("qpos0", _C.c_void_p), ("reuse_data", _C.c_int32), ("max_reuse_data", _C.c_int32),
                ("seed_reuse", _C.c_uint64), ("discrete_action", _C.c_int32), ("ac_space_normal", _C.c_int32),
                ("simple_planner_range", _C.c_double), ("simple_max_iter", _C.c_int32), ("debug_block_mod", _C.c_int32),
                ("use_ik_target", _C.c_int32), ("ik_body", _C.c_int32), ("ik_site_local", _C.c_double * 3),
                ("ik_world_lo", _C.c_double * 3), ("ik_world_hi", _C.c_double * 3), ("ik_max_steps", _C.c_int32), ("ik_pad_", _C.c_int32),
                ("ik_tol", _C.c_double)]


class NativeMoPARolloutRunner:
    """The collection loop with every step of a tick as a CUDA kernel of
    libmopa_b200 (csrc/rollout.cu): no host round trip inside a tick, the policy is evaluated once per
    tick on the observations of ALL environments (its output is used where a macro action starts).

    ``policy(obs [n,40] f32, env_gid [n] i64, macro_index [n] i64) -> [n,7]`` actions in [-1, 1]; with
    ``config.discrete_action`` it returns ``(actions [n,7], ac_type [n])`` (1 = motion planner, 0 = direct execution).
    """

    def __init__(self, venv, config=None, policy=None, transition_capacity=1 << 20, rrt_capacity=1024):
        import torch

        from .capi import check, lib

        self.torch, self.venv, self.cfg = torch, venv, config or MoPAConfig()
        cfg, m, dev = self.cfg, venv.model, venv.dev
        self.dev = dev
        ignored, passive, ref = env_planner_inputs(type(venv), m)
        self.planner = NativePlanner(m, passive, ignored, cfg.contact_threshold, cfg.range, 0.005, cfg.seed, venv.device_index)
        self.action_dim = int(getattr(venv, "ACTION_DIM", 7))   # 8 for the lift task (7 joint entries + gripper)
        self.policy = policy or UniformPolicy(torch, dev, cfg.seed + 17 * int(venv.env_ids[0]), self.action_dim)
        n = venv.n
        self.n = n
        self.env_gid = torch.as_tensor(venv.env_ids, dtype=torch.int64, device=dev)
        self.macro_index = torch.zeros(n, dtype=torch.int64, device=dev)
        self.slab = torch.zeros(n, TRANSITION_FLOATS, dtype=torch.float32, device=dev)
        self.emit_flag = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.transitions = torch.zeros(transition_capacity, TRANSITION_FLOATS, dtype=torch.float32, device=dev)
        self._counters = torch.zeros(24, dtype=torch.int64, device=dev)
        self.ep_stats = torch.zeros(n, 5, dtype=torch.float64, device=dev)   # per env: episodes, sum len, sum rew, sum success, sum contact force
        if cfg.reuse_data and not 1 <= int(cfg.max_reuse_data) <= 32:
            raise NotImplementedError("max_reuse_data = %r: the relabelling kernel keeps up to 32 (start, goal) pairs per plan" % (cfg.max_reuse_data,))
        if cfg.action_range / (0.8 * cfg.ac_scale) > 16:
            raise NotImplementedError("action_range / (0.8 * ac_scale) = %.1f: the straight-line planner checks at most 16 interpolation points"
                                      % (cfg.action_range / (0.8 * cfg.ac_scale)))
        self.max_reuse = int(cfg.max_reuse_data)
        jid = [list(m.jnt_qposadr).index(a) for a in ref]
        c = _RolloutConfig()
        c.n_envs, c.max_iter, c.max_path, c.max_traj, c.rrt_capacity = n, cfg.max_iter, cfg.max_path, cfg.max_traj, min(rrt_capacity, max(n, 16))
        c.num_trials, c.invalid_target_handling, c.interpolation = cfg.num_trials, int(cfg.invalid_target_handling), int(cfg.interpolation)
        c.omega, c.action_range, c.ac_scale, c.discount = cfg.omega, cfg.action_range, cfg.ac_scale, cfg.discount_factor
        c.step_size, c.joint_margin, c.range = cfg.step_size, cfg.joint_margin, cfg.range
        c.seed_env, c.env_id_offset = venv.seed, int(venv.env_ids[0])
        for k in range(len(ref)):   # unlimited hinges (Pusher joint0): infinite range = never clipped, SO(2) in the planner
            lim = bool(m.jnt_limited[jid[k]])
            c.jnt_lo[k] = float(m.jnt_range[jid[k], 0]) if lim else -np.inf
            c.jnt_hi[k] = float(m.jnt_range[jid[k], 1]) if lim else np.inf
            c.init_qpos[k] = float(venv.INIT_QPOS[k])
        self._qpos0 = np.ascontiguousarray(m.qpos0, dtype=np.float64)
        c.qpos0 = self._qpos0.ctypes.data
        c.reuse_data, c.max_reuse_data, c.seed_reuse = int(cfg.reuse_data), self.max_reuse, (int(cfg.seed) + 0x5EED) & 0xFFFFFFFFFFFFFFFF
        c.discrete_action = int(cfg.discrete_action)
        if cfg.ac_space_type not in ("piecewise", "normal"):
            raise NotImplementedError("ac_space_type %r" % (cfg.ac_space_type,))   # rl/sac_agent.py:174-175 raises as well
        c.ac_space_normal = int(cfg.ac_space_type == "normal")
        c.simple_planner_range, c.simple_max_iter = float(cfg.simple_planner_range), int(cfg.simple_max_iter)
        c.debug_block_mod = int(cfg.debug_block_mod)
        c.use_ik_target = int(bool(cfg.use_ik_target))
        if cfg.use_ik_target:
            from .inverse_kinematics import site_frame

            if cfg.discrete_action or self.action_dim != 8:
                raise NotImplementedError("use_ik_target: built for the 8-entry Cartesian action (default[3], quat[4], gripper) of the lift "
                                          "task without discrete_action (rl/trainer.py:113-135)")
            sid = m.site_name2id(cfg.ik_target)
            if not np.allclose(m.site_quat[sid], [1, 0, 0, 0]):
                raise NotImplementedError("use_ik_target: the ik_target site must not be rotated against its body")
            body, local = site_frame(m, venv.dyn, cfg.ik_target)
            c.ik_body = int(body)
            lo, hi = getattr(venv, "WORLD", ((-1.2, -1.2, 0.0), (1.2, 1.2, 2.0)))   # SawyerEnv.min_world_size / max_world_size (sawyer.py:52-53)
            for k in range(3):
                c.ik_site_local[k], c.ik_world_lo[k], c.ik_world_hi[k] = float(local[k]), float(lo[k]), float(hi[k])
            c.ik_max_steps, c.ik_tol = int(cfg.ik_max_steps), float(cfg.ik_tol)
        L = lib()
        L.mopa_rollout_create.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p,
                                          _C.c_void_p, _C.c_int64, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_int32, _C.c_void_p, _C.POINTER(_C.c_void_p)]
        L.mopa_rollout_destroy.argtypes = [_C.c_void_p]
        L.mopa_rollout_destroy.restype = None
        L.mopa_rollout_pre.argtypes = [_C.c_void_p, _C.c_int32, _C.c_void_p]
        L.mopa_rollout_step.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p]
        L.mopa_rollout_step_discrete.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p]
        L.mopa_rollout_step_ik.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p]
        L.mopa_rollout_busy.argtypes = [_C.c_void_p]
        L.mopa_rollout_launches.argtypes = [_C.c_void_p]
        L.mopa_rollout_launches.restype = _C.c_int64
        L.mopa_rollout_env_ms.argtypes = [_C.c_void_p, _C.c_int32, _C.POINTER(_C.c_double)]
        L.mopa_rollout_pack.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_int32, _C.c_void_p]
        self._L, self._check = L, check
        venv.reset()
        h = _C.c_void_p()
        check(L.mopa_rollout_create(venv.h, self.planner.h, _C.byref(venv.buf), _C.byref(c), self.macro_index.data_ptr(), self.slab.data_ptr(),
                                    self.emit_flag.data_ptr(), self.transitions.data_ptr(), transition_capacity, self._counters.data_ptr(),
                                    None, None, 0, self.ep_stats.data_ptr(),
                                    _C.byref(h)))
        self.h = h
        self.ticks = 0
        self.last_emitted = None         # (slab [n,92], emit_flag [n]) of the latest tick: the main records, dense by environment

    def close(self):
        if getattr(self, "h", None):
            self._L.mopa_rollout_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return _C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def tick(self, wait_rrt=False):
        self._check(self._L.mopa_rollout_pre(self.h, int(wait_rrt), self._stream()))
        ac = self.policy(self.venv.obs, self.env_gid, self.macro_index)
        if self.cfg.discrete_action:   # policy -> (ac["default"] [n,7], ac["ac_type"] [n])
            ac, ac_type = ac
            ac_type = ac_type.to(device=self.dev, dtype=self.torch.uint8).contiguous()
        ac = ac.to(device=self.dev, dtype=self.torch.float32).contiguous()
        if tuple(ac.shape) != (self.n, self.action_dim):
            raise ValueError("policy returned actions of shape %s, expected (%d, %d)" % (tuple(ac.shape), self.n, self.action_dim))
        if self.cfg.discrete_action:
            self._check(self._L.mopa_rollout_step_discrete(self.h, ac.data_ptr(), ac_type.data_ptr(), self._stream()))
            self._keep = (ac, ac_type)
        elif self.cfg.use_ik_target:   # policy -> (default[3], quat[4], gripper): the IK solve runs inside the step
            self._check(self._L.mopa_rollout_step_ik(self.h, ac.data_ptr(), self._stream()))
            self._keep = ac
        else:
            self._check(self._L.mopa_rollout_step(self.h, ac.data_ptr(), self._stream()))
            self._keep = ac
        self.ticks += 1
        self.last_emitted = (self.slab, self.emit_flag)

    def pack(self, send, capacity):
        """Step 1 of the replay exchange (mopa_rollout_pack): the records (main and relabelled) emitted since the previous call, at
        most ``capacity`` of them, compact into ``send`` [1 + capacity, 92] behind a header row; the rest stays queued."""
        self._check(self._L.mopa_rollout_pack(self.h, send.data_ptr(), int(capacity), self._stream()))

    def drain(self, max_ticks=64):
        """Tick until no environment waits for an RRT plan (end of a collection run)."""
        for _ in range(max_ticks):
            self.tick(wait_rrt=True)
            if self.counter_values()["waiting"] == 0:
                break

    def episode_stats(self):
        """Means over the finished episodes of all environments - what MoPARolloutRunner.run_episode reports per episode
        (rl/mopa_rollouts.py:662-681): len, rew, episode_success, contact_force."""
        s = self.ep_stats.sum(dim=0).cpu().numpy()
        k = max(s[0], 1.0)
        return dict(episodes=int(s[0]), len=s[1] / k, rew=s[2] / k, episode_success=s[3] / k, contact_force=s[4] / k)

    def counter_values(self):
        v = self._counters.cpu().numpy()
        return {k: int(v[i]) for i, k in enumerate(COUNTER_NAMES)}

    @property
    def counters(self):
        return self.counter_values()

    @property
    def env_steps(self):
        return self.counter_values()["env_steps"]

    @property
    def n_transitions(self):
        return self.counter_values()["transitions"]

    @property
    def launches(self):
        return int(self._L.mopa_rollout_launches(self.h))

    def rrt_stats(self):
        """(last batch ms, batches finished, mean ms per batch, mean ticks from launch to finalisation)."""
        out = (_C.c_double * 4)()
        self._L.mopa_rollout_rrt_stats.argtypes = [_C.c_void_p, _C.c_void_p]
        self._check(self._L.mopa_rollout_rrt_stats(self.h, out))
        return tuple(out)

    def env_kernel_ms(self, n_last):
        """Mean device time of the env-step kernel over the latest ``n_last`` ticks (synchronises)."""
        out = _C.c_double()
        self._check(self._L.mopa_rollout_env_ms(self.h, int(n_last), _C.byref(out)))
        return out.value


def run_episodes(venv, config=None, policy=None, episodes_per_env=1, max_ticks=100000):
    """Evaluation path (MoPARolloutRunner.run_episode, rl/mopa_rollouts.py:401-681, batched): run the collection loop
    with the given (deterministic) policy until every environment has finished ``episodes_per_env`` episodes and return
    the per-episode means the reference logs (len, rew, episode_success, contact_force) plus the planner counters."""
    runner = NativeMoPARolloutRunner(venv, config, policy=policy, transition_capacity=1 << 16)
    for t in range(max_ticks):
        runner.tick()
        if t % 16 == 15 and float(runner.ep_stats[:, 0].min()) >= episodes_per_env:
            break
    info = runner.episode_stats()
    info.update({k: v for k, v in runner.counters.items() if k in ("mp", "rl", "interpolation", "mp_fail", "approximate", "invalid")})
    info["per_env"] = runner.ep_stats.cpu().numpy().copy()   # [n, 5]: episodes, sum of len / rew / success / contact force
    runner.close()
    return info
