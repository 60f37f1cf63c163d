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
    max_path = 384
    max_traj = 1280
    seed = 1234
    reuse_data = False         # scripts/3d/push/mopa.sh sets True: relabel (start, goal) pairs of every executed plan
    max_reuse_data = 15
    ac_space_type = "piecewise"  # "normal" (lift / assembly / 2d mopa_discrete.sh): displacement = a * action_range
    discrete_action = False    # scripts/3d/*/mopa_discrete.sh (with omega = 0): the policy's ac_type picks planner / direct

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
    """planner_inputs for a vectorised env class (STATIC_BODIES / MANIPULATION_BODIES)."""
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


class VecMoPARolloutRunner:
    def __init__(self, venv, config=None, policy=None, transition_capacity=1 << 20):
        import torch

        if config is not None and (config.discrete_action or config.reuse_data or config.ac_space_type != "piecewise"):
            raise NotImplementedError("discrete_action / reuse_data / ac_space_type 'normal' are implemented by NativeMoPARolloutRunner only")

        self.torch = torch
        self.venv, self.cfg = venv, config or MoPAConfig()
        cfg, m, dev = self.cfg, venv.model, venv.dev
        self.dev = dev
        ignored, passive, ref = planner_inputs(m)
        assert ref == list(range(7)), "arm joints are expected to be the first qpos entries (sac_agent.py uses [:7])"
        self.planner = NativePlanner(m, passive, ignored, cfg.contact_threshold, cfg.range, 0.005, cfg.seed, venv.device_index)
        self.policy = policy or UniformPolicy(torch, dev, cfg.seed + 17 * int(venv.env_ids[0]))
        n = venv.n
        self.nq = m.nq
        self.row = ((m.nq + 3) // 4) * 4
        jid = [list(m.jnt_qposadr).index(a) for a in ref]
        self.jlo = torch.as_tensor(m.jnt_range[jid, 0], device=dev)
        self.jhi = torch.as_tensor(m.jnt_range[jid, 1], device=dev)
        f64 = torch.float64
        self.traj = torch.zeros(n, cfg.max_traj, 7, dtype=f64, device=dev)
        self.traj_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.traj_pos = torch.zeros(n, dtype=torch.int32, device=dev)
        self.kind = torch.zeros(n, dtype=torch.uint8, device=dev)      # 0 direct, 1 plan, 2 planner failure
        self.pending = torch.zeros(n, dtype=torch.bool, device=dev)    # a macro action is in flight
        self.prev_ob = torch.zeros(n, 40, dtype=torch.float32, device=dev)
        self.ac = torch.zeros(n, 8, dtype=torch.float32, device=dev)
        self.meta_rew = torch.zeros(n, dtype=f64, device=dev)
        self.executed = torch.zeros(n, dtype=torch.int32, device=dev)
        self.macro_done = torch.zeros(n, dtype=torch.bool, device=dev)
        self.step_action = torch.zeros(n, 8, dtype=torch.float32, device=dev)
        self.step_mode = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.plan_calls = 0
        self.plan_count = torch.zeros(n, dtype=torch.int64, device=dev)
        self.macro_index = torch.zeros(n, dtype=torch.int64, device=dev)
        self.env_gid = torch.as_tensor(venv.env_ids, dtype=torch.int64, device=dev)
        self.transitions = torch.zeros(transition_capacity, TRANSITION_FLOATS, dtype=torch.float32, device=dev)
        self.n_transitions = 0
        self.env_steps = 0
        self.counters = dict(mp=0, rl=0, interpolation=0, mp_fail=0, approximate=0, invalid=0, densify_fallback=0, episodes=0,
                             success=0, mp_path_len=0, interpolation_path_len=0)
        self.launches = 0
        self.plan_stream = torch.cuda.Stream(device=dev)
        self.rrt_queue, self.rrt_inflight = [], None
        self.n_waiting = 0
        self.async_rrt = True
        self.step_events = None      # set to [] to collect CUDA events around every env-step launch
        self.last_emitted = None     # transition records emitted by the latest tick (for the replay exchange)
        venv.reset()

    # ---------------------------------------------------------------- batched planner glue
    def _valid(self, states64):
        """states64 [M, nq] float64 -> bool [M] (MujocoStateValidityChecker::isValid, batched)."""
        torch = self.torch
        M = states64.shape[0]
        if M == 0:
            return torch.zeros(0, dtype=torch.bool, device=self.dev)
        q = torch.zeros(M, self.row, dtype=torch.float32, device=self.dev)
        q[:, :self.nq] = states64.float()
        out = torch.zeros(M, dtype=torch.int32, device=self.dev)
        self.planner.is_valid_device(q.data_ptr(), self.row, M, out.data_ptr(), 0, torch.cuda.current_stream(self.dev).cuda_stream)
        self.launches += 1
        self._keep_v = (q, out)
        return (out & 1).bool()

    def _clip_qpos(self, q):
        """SACAgent.clip_qpos (rl/sac_agent.py:237-259): only when some limited joint is out of range."""
        torch = self.torch
        arm = q[:, :7]
        out = ((arm < self.jlo) | (arm > self.jhi)).any(dim=1)
        clipped = torch.minimum(torch.maximum(arm, self.jlo + self.cfg.joint_margin), self.jhi - self.cfg.joint_margin)
        q = q.clone()
        q[:, :7] = torch.where(out[:, None], clipped, arm)
        return q

    def _interp_points(self, start, target, jmax):
        """simple_interpolate (rl/sac_agent.py:262-298): per-joint steps of at most 0.8*ac_scale.
        Returns nstep [M] and the interpolated states [M, jmax, nq] (rows j >= nstep are padding)."""
        torch = self.torch
        lim = self.cfg.ac_scale * 0.8
        diff = target[:, :7] - start[:, :7]
        sf = torch.clamp((diff.abs() / lim).max(dim=1).values, min=1.0)
        # scales only count joints whose |diff| exceeds the limit; below it the factor is 1 anyway
        nstep = torch.clamp(sf.floor().to(torch.int64), max=jmax)
        scaled = diff / sf[:, None]
        pts = start[:, None, :].repeat(1, jmax, 1)
        run = start[:, :7].clone()
        for j in range(jmax):  # the reference accumulates interp_qpos += scaled_ac: same running sum
            run = run + scaled
            pts[:, j, :7] = run
        return nstep, pts

    def _plan(self, idx, ac):
        """Plan for envs `idx` (int64 tensor) with actions ac [K,7].  Fills traj / traj_len / kind."""
        torch, cfg, venv = self.torch, self.cfg, self.venv
        K = idx.numel()
        curr = venv.qpos[idx]
        a = ac.double()
        w = cfg.omega
        disp = torch.where(a.abs() < w, a / (w / cfg.ac_scale),
                           torch.sign(a) * (cfg.ac_scale + (cfg.action_range - cfg.ac_scale) * ((a.abs() - w) / (1 - w))))
        target = curr.clone()
        target[:, :7] = torch.minimum(torch.maximum(curr[:, :7] + disp, self.jlo), self.jhi)
        ok = self._valid(target)
        if cfg.invalid_target_handling and not bool(ok.all()):
            bad = torch.nonzero(~ok).squeeze(1)
            t = target[bad].clone()
            c = curr[bad]
            cands = torch.empty(bad.numel(), cfg.num_trials, self.nq, dtype=torch.float64, device=self.dev)
            for k in range(cfg.num_trials):
                d = c - t
                t = t + cfg.step_size * d / torch.linalg.norm(d, dim=1, keepdim=True)
                cands[:, k] = t
            v = self._valid(cands.reshape(-1, self.nq)).reshape(bad.numel(), cfg.num_trials)
            anyv = v.any(dim=1)
            first = torch.argmax(v.to(torch.int8), dim=1)
            chosen = cands[torch.arange(bad.numel(), device=self.dev), first]
            target[bad] = torch.where(anyv[:, None], chosen, cands[:, -1])
            ok[bad] = anyv
        kind = torch.full((K,), 2, dtype=torch.uint8, device=self.dev)   # failure unless proven otherwise
        tlen = torch.ones(K, dtype=torch.int32, device=self.dev)
        n_invalid = int((~ok).sum())
        self.counters["invalid"] += n_invalid
        good = torch.nonzero(ok).squeeze(1)
        if good.numel():
            c = self._clip_qpos(curr[good])
            tg = target[good]
            jmax = 16
            nstep, pts = self._interp_points(c, tg, jmax)
            jj = torch.arange(jmax, device=self.dev)
            live = jj[None, :] < nstep[:, None]
            v = self._valid(pts.reshape(-1, self.nq)).reshape(-1, jmax)
            straight = (v | ~live).all(dim=1)
            # interpolation success: the interpolated states followed by the target
            s_idx = torch.nonzero(straight).squeeze(1)
            if s_idx.numel():
                rows = idx[good[s_idx]]
                tr = torch.zeros(s_idx.numel(), jmax + 1, 7, dtype=torch.float64, device=self.dev)
                tr[:, :jmax] = pts[s_idx][:, :, :7]
                tr[torch.arange(s_idx.numel(), device=self.dev), nstep[s_idx]] = tg[s_idx][:, :7]
                self.traj[rows, :jmax + 1] = tr
                kind[good[s_idx]] = 1
                tlen[good[s_idx]] = (nstep[s_idx] + 1).to(torch.int32)
                self.counters["interpolation"] += int(s_idx.numel())
                self.counters["interpolation_path_len"] += int((nstep[s_idx] + 1).sum())
            r_idx = torch.nonzero(~straight).squeeze(1)
        else:
            r_idx = good
        self.counters["mp_fail"] += int((~ok).sum())
        self.kind[idx] = kind
        self.traj_len[idx] = tlen
        self.traj_pos[idx] = 0
        if r_idx.numel():
            # straight line blocked: RRT-Connect, asynchronously (these envs wait, kind 3)
            self.rrt_queue.append((idx[good[r_idx]], c[r_idx], tg[r_idx]))

    def _rrt_launch(self, rows, start, goal):
        """Enqueue RRT-Connect for env rows `rows` on the planner stream; the envs wait (kind 3) until
        the batch is finalised by a later tick, so the planner runs under the env-step kernels."""
        torch, cfg = self.torch, self.cfg
        R = rows.numel()
        row, mp = self.row, cfg.max_path
        s32 = torch.zeros(R, row, dtype=torch.float32, device=self.dev)
        g32 = torch.zeros(R, row, dtype=torch.float32, device=self.dev)
        s32[:, :self.nq] = start.float()
        g32[:, :self.nq] = goal.float()
        keys = (self.env_gid[rows] << 32) + self.plan_count[rows]   # invariant to batching / GPU count
        self.plan_count[rows] += 1
        self.plan_calls += R
        path = torch.zeros(R, mp, row, dtype=torch.float32, device=self.dev)
        ids = torch.zeros(R, mp, dtype=torch.int32, device=self.dev)
        plen = torch.zeros(R, dtype=torch.int32, device=self.dev)
        status = torch.zeros(R, dtype=torch.int32, device=self.dev)
        main = torch.cuda.current_stream(self.dev)
        ready = torch.cuda.Event()
        ready.record(main)
        self.plan_stream.wait_event(ready)
        self.planner.plan_device(s32.data_ptr(), g32.data_ptr(), row, keys.data_ptr(), R, cfg.max_iter, path.data_ptr(), ids.data_ptr(),
                                 mp, plen.data_ptr(), status.data_ptr(), 0, 0, self.plan_stream.cuda_stream)
        done = torch.cuda.Event()
        done.record(self.plan_stream)
        self.launches += 1
        self.kind[rows] = 3
        self.traj_len[rows] = 1
        self.traj_pos[rows] = 0
        return dict(rows=rows, start=start, s32=s32, g32=g32, keys=keys, path=path, ids=ids, plen=plen, status=status, done=done)

    def _rrt_finalize(self, batch):
        """Re-base, densify and store the paths of a finished RRT batch (SamplingBasedPlanner.plan /
        PlannerAgent.plan / SACAgent.plan densification, rl/sac_agent.py:216-233)."""
        torch, cfg = self.torch, self.cfg
        torch.cuda.current_stream(self.dev).wait_event(batch["done"])
        rows_all, start, path, plen, status = batch["rows"], batch["start"], batch["path"], batch["plen"], batch["status"]
        mp = cfg.max_path
        R = rows_all.numel()
        kind = torch.full((R,), 2, dtype=torch.uint8, device=self.dev)
        tlen = torch.ones(R, dtype=torch.int32, device=self.dev)
        okp = status == 0
        n_fail = int((~okp).sum())
        self.counters["approximate"] += n_fail
        self.counters["mp_fail"] += n_fail
        w = torch.nonzero(okp).squeeze(1)
        if w.numel():
            self._densify(rows_all, w, start, path, plen, kind, tlen)
        self.kind[rows_all] = kind
        self.traj_len[rows_all] = tlen
        self.traj_pos[rows_all] = 0

    def _densify(self, rows_all, w, start, path, plen, kind, tlen):
        torch, cfg = self.torch, self.cfg
        mp = cfg.max_path
        self.counters["mp"] += int(w.numel())
        P = path[w][:, :, :7].double()
        L = plen[w].to(torch.int64)                                  # rows incl. the start row
        st = start[w][:, :7]
        # SamplingBasedPlanner.plan re-bases the path on `start`; PlannerAgent.plan drops the first row
        P = st[:, None, :] + (P - P[:, :1])
        hops_end = P[:, 1:]                                          # traj[i]
        hops_start = torch.cat([st[:, None, :], P[:, 1:-1]], dim=1)  # start of hop i
        H = mp - 1
        hop_live = torch.arange(H, device=self.dev)[None, :] < (L - 1)[:, None]
        diff = hops_end - hops_start
        if cfg.interpolation:
            lim = cfg.ac_scale * 0.8
            need = (diff.abs() > cfg.ac_scale).any(dim=2) & hop_live
            sf = torch.clamp((diff.abs() / lim).max(dim=2).values, min=1.0)
            kmax = int(cfg.range / lim) + 1
            nst = torch.where(need, torch.clamp(sf.floor().to(torch.int64), max=kmax), torch.zeros_like(L)[:, None].expand(-1, H))
            scaled = diff / sf[..., None]
            inter = torch.empty(w.numel(), H, kmax, 7, dtype=torch.float64, device=self.dev)
            run = hops_start.clone()
            for j in range(kmax):
                run = run + scaled
                inter[:, :, j] = run
            live = (torch.arange(kmax, device=self.dev)[None, None, :] < nst[..., None])
            if bool(live.any()):
                full = start[w][:, None, None, :].expand(-1, H, kmax, -1).clone()
                full[..., :7] = inter
                sel = torch.nonzero(live.reshape(-1)).squeeze(1)
                vv = torch.ones(live.numel(), dtype=torch.bool, device=self.dev)
                vv[sel] = self._valid(full.reshape(-1, self.nq)[sel])
                hop_bad = (~vv.reshape(live.shape) & live).any(dim=2)
                if bool(hop_bad.any()):
                    # the reference would call the simple planner / main planner here (sac_agent.py:300-311);
                    # such hops keep only their end point and are counted
                    self.counters["densify_fallback"] += int(hop_bad.sum())
                    nst = torch.where(hop_bad, torch.zeros_like(nst), nst)
                    live = live & ~hop_bad[..., None]
        else:
            kmax = 1
            nst = torch.zeros(w.numel(), H, dtype=torch.int64, device=self.dev)
            inter = hops_end[:, :, None, :]
            live = torch.zeros(w.numel(), H, 1, dtype=torch.bool, device=self.dev)
        cnt = (nst + 1) * hop_live
        off = torch.cumsum(cnt, dim=1) - cnt
        total = cnt.sum(dim=1)
        over = total > cfg.max_traj
        rows = rows_all[w]
        out = torch.zeros(w.numel(), cfg.max_traj, 7, dtype=torch.float64, device=self.dev)
        ar = torch.arange(w.numel(), device=self.dev)[:, None].expand(-1, H)
        for j in range(kmax):
            msk = live[:, :, j] & hop_live & ~over[:, None]
            if bool(msk.any()):
                out[ar[msk], (off + j)[msk]] = inter[:, :, j][msk]
        msk = hop_live & ~over[:, None]
        out[ar[msk], (off + nst)[msk]] = hops_end[msk]
        self.traj[rows] = out
        kind[w] = torch.where(over, torch.full_like(total, 2), torch.ones_like(total)).to(torch.uint8)
        tlen[w] = torch.where(over, torch.ones_like(total), total).to(torch.int32)
        self.counters["mp_fail"] += int(over.sum())
        self.counters["mp_path_len"] += int(total[~over].sum())

    # ---------------------------------------------------------------- one env.step for every env
    def tick(self):
        torch, cfg, venv = self.torch, self.cfg, self.venv
        if self.rrt_inflight is not None and (not self.async_rrt or self.rrt_inflight["done"].query()):
            self._rrt_finalize(self.rrt_inflight)
            self.rrt_inflight = None
        need = torch.nonzero(self.traj_pos >= self.traj_len).squeeze(1)
        self.last_emitted = None
        if need.numel():
            # finished macro actions -> transition records; finished episodes -> reset
            fin = need[self.pending[need]]
            if fin.numel():
                rec = torch.zeros(fin.numel(), TRANSITION_FLOATS, dtype=torch.float32, device=self.dev)
                rec[:, 0:40] = self.prev_ob[fin]
                rec[:, 40:48] = self.ac[fin]
                rec[:, 48] = self.meta_rew[fin].float()
                rec[:, 49] = self.macro_done[fin].float()
                rec[:, 50] = (self.executed[fin] - 1).clamp(min=0).float()
                rec[:, 51] = self.env_gid[fin].float()
                rec[:, 52:92] = venv.obs[fin]
                k = fin.numel()
                w0 = self.n_transitions % self.transitions.shape[0]
                k1 = min(k, self.transitions.shape[0] - w0)
                self.transitions[w0:w0 + k1] = rec[:k1]
                if k1 < k:
                    self.transitions[:k - k1] = rec[k1:]
                self.n_transitions += k
                self.last_emitted = rec
                dn = fin[self.macro_done[fin]]
                if dn.numel():
                    self.counters["episodes"] += int(dn.numel())
                    self.counters["success"] += int(venv.success[dn].sum())
                    venv.reset(dn.cpu().numpy())
            venv.has_prev[need] = 0                                   # env._reset_prev_state()
            obs = venv.obs[need]
            ac = self.policy(obs, self.env_gid[need], self.macro_index[need]).float().clamp(-1, 1)
            self.macro_index[need] += 1
            self.prev_ob[need] = obs
            self.ac[need, :7] = ac
            self.meta_rew[need] = 0
            self.executed[need] = 0
            self.macro_done[need] = False
            self.pending[need] = True
            is_mp = (ac.abs() > cfg.omega).any(dim=1)
            d_idx = need[~is_mp]
            if d_idx.numel():
                self.kind[d_idx] = 0
                self.traj_len[d_idx] = 1
                self.traj_pos[d_idx] = 0
                self.counters["rl"] += int(d_idx.numel())
            p_idx = need[is_mp]
            if p_idx.numel():
                self._plan(p_idx, ac[is_mp])
        if self.rrt_inflight is None and self.rrt_queue:
            rows = torch.cat([r for r, _, _ in self.rrt_queue])
            st = torch.cat([a for _, a, _ in self.rrt_queue])
            gl = torch.cat([b for _, _, b in self.rrt_queue])
            self.rrt_queue = []
            self.rrt_inflight = self._rrt_launch(rows, st, gl)
            if not self.async_rrt:
                self._rrt_finalize(self.rrt_inflight)
                self.rrt_inflight = None
        elif self.rrt_queue:
            for r, _, _ in self.rrt_queue:      # queued behind the batch in flight: wait as well
                self.kind[r] = 3
                self.traj_len[r] = 1
                self.traj_pos[r] = 0
        self.n_waiting = int((self.kind == 3).sum()) if (self.rrt_inflight is not None or self.rrt_queue) else 0
        # stage the action of every env for this tick
        kind = self.kind
        self.step_mode.copy_(kind)
        direct = kind == 0
        self.step_action[:, :7] = torch.where(direct[:, None], (self.ac[:, :7].double() / cfg.omega).float(), self.step_action[:, :7])
        plan = kind == 1
        if bool(plan.any()):
            pos = self.traj_pos.to(torch.int64).clamp(max=cfg.max_traj - 1)
            nxt = self.traj[torch.arange(venv.n, device=self.dev), pos]
            delta = (nxt - venv.qpos[:, :7]).float()                  # env.form_action(next_qpos)
            self.step_action[:, :7] = torch.where(plan[:, None], delta, self.step_action[:, :7])
        stepping = kind != 3                                          # envs waiting for their RRT plan do not step
        mask = stepping.to(torch.uint8) if self.n_waiting else None
        if self.step_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            venv.step(self.step_action, self.step_mode, mask)
            e1.record()
            self.step_events.append((e0, e1))
        else:
            venv.step(self.step_action, self.step_mode, mask)
        self.launches += 1
        disc = torch.pow(torch.full_like(self.meta_rew, cfg.discount_factor), self.traj_pos.double())
        gain = torch.where(plan, disc, torch.ones_like(disc)) * venv.reward
        self.meta_rew += torch.where(stepping, gain, torch.zeros_like(gain))
        one = stepping.to(torch.int32)
        self.executed += one
        self.traj_pos += one
        done = venv.done.bool() & stepping
        self.macro_done |= done
        self.traj_len = torch.where(done, self.traj_pos, self.traj_len)
        stepped = venv.n - self.n_waiting
        self.env_steps += stepped
        return stepped

    def drain(self):
        """Finish any RRT batch in flight (end of a collection run)."""
        while self.rrt_inflight is not None or self.rrt_queue:
            if self.rrt_inflight is not None:
                self.rrt_inflight["done"].synchronize()
            self.tick()


# ------------------------------------------------------------------------------------ native runner
import ctypes as _C  # noqa: E402

COUNTER_NAMES = ("mp", "rl", "interpolation", "mp_fail", "approximate", "invalid", "densify_fallback", "episodes", "success",
                 "mp_path_len", "interpolation_path_len", "env_steps", "transitions", "rrt_dropped", "rrt_problems", "waiting", "reused")


class _RolloutConfig(_C.Structure):
    _fields_ = [(k, _C.c_int32) for k in ("n_envs", "max_iter", "max_path", "max_traj", "rrt_capacity", "num_trials",
                                          "invalid_target_handling", "interpolation")] + \
               [(k, _C.c_double) for k in ("omega", "action_range", "ac_scale", "discount", "step_size", "joint_margin", "range")] + \
               [("seed_env", _C.c_uint64), ("env_id_offset", _C.c_int64), ("jnt_lo", _C.c_double * 7), ("jnt_hi", _C.c_double * 7),
                ("init_qpos", _C.c_double * 7), ("qpos0", _C.c_void_p), ("reuse_data", _C.c_int32), ("max_reuse_data", _C.c_int32),
                ("seed_reuse", _C.c_uint64), ("discrete_action", _C.c_int32), ("ac_space_normal", _C.c_int32)]


class NativeMoPARolloutRunner:
    """Same collection loop as ``VecMoPARolloutRunner`` with every step of a tick as a CUDA kernel of
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
        self._counters = torch.zeros(18, dtype=torch.int64, device=dev)
        self.ep_stats = torch.zeros(n, 5, dtype=torch.float64, device=dev)   # per env: episodes, sum len, sum rew, sum success, sum contact force
        self.max_reuse = max(1, min(16, int(cfg.max_reuse_data)))
        self.reuse_capacity = 2 * n      # relabelled records per tick that take part in the replay exchange (steady state: ~0.4 n)
        self.reuse_slab = torch.zeros(self.reuse_capacity, TRANSITION_FLOATS, dtype=torch.float32, device=dev) if cfg.reuse_data else None
        self.reuse_count = torch.zeros(1, dtype=torch.int32, device=dev) if cfg.reuse_data else None
        self._reuse_iota = torch.arange(self.reuse_capacity, dtype=torch.int32, device=dev)
        jid = [list(m.jnt_qposadr).index(a) for a in ref]
        c = _RolloutConfig()
        c.n_envs, c.max_iter, c.max_path, c.max_traj, c.rrt_capacity = n, cfg.max_iter, cfg.max_path, cfg.max_traj, min(rrt_capacity, max(n, 16))
        c.num_trials, c.invalid_target_handling, c.interpolation = cfg.num_trials, int(cfg.invalid_target_handling), int(cfg.interpolation)
        c.omega, c.action_range, c.ac_scale, c.discount = cfg.omega, cfg.action_range, cfg.ac_scale, cfg.discount_factor
        c.step_size, c.joint_margin, c.range = cfg.step_size, cfg.joint_margin, cfg.range
        c.seed_env, c.env_id_offset = venv.seed, int(venv.env_ids[0])
        for k in range(7):
            c.jnt_lo[k], c.jnt_hi[k], c.init_qpos[k] = float(m.jnt_range[jid[k], 0]), float(m.jnt_range[jid[k], 1]), float(venv.INIT_QPOS[k])
        self._qpos0 = np.ascontiguousarray(m.qpos0, dtype=np.float64)
        c.qpos0 = self._qpos0.ctypes.data
        c.reuse_data, c.max_reuse_data, c.seed_reuse = int(cfg.reuse_data), self.max_reuse, (int(cfg.seed) + 0x5EED) & 0xFFFFFFFFFFFFFFFF
        c.discrete_action = int(cfg.discrete_action)
        if cfg.ac_space_type not in ("piecewise", "normal"):
            raise NotImplementedError("ac_space_type %r" % (cfg.ac_space_type,))   # rl/sac_agent.py:174-175 raises as well
        c.ac_space_normal = int(cfg.ac_space_type == "normal")
        L = lib()
        L.mopa_rollout_create.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p,
                                          _C.c_void_p, _C.c_int64, _C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_int32, _C.c_void_p, _C.POINTER(_C.c_void_p)]
        L.mopa_rollout_destroy.argtypes = [_C.c_void_p]
        L.mopa_rollout_destroy.restype = None
        L.mopa_rollout_pre.argtypes = [_C.c_void_p, _C.c_int32, _C.c_void_p]
        L.mopa_rollout_step.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p]
        L.mopa_rollout_step_discrete.argtypes = [_C.c_void_p, _C.c_void_p, _C.c_void_p, _C.c_void_p]
        L.mopa_rollout_busy.argtypes = [_C.c_void_p]
        L.mopa_rollout_launches.argtypes = [_C.c_void_p]
        L.mopa_rollout_launches.restype = _C.c_int64
        L.mopa_rollout_env_ms.argtypes = [_C.c_void_p, _C.c_int32, _C.POINTER(_C.c_double)]
        self._L, self._check = L, check
        venv.reset()
        h = _C.c_void_p()
        check(L.mopa_rollout_create(venv.h, self.planner.h, _C.byref(venv.buf), _C.byref(c), self.macro_index.data_ptr(), self.slab.data_ptr(),
                                    self.emit_flag.data_ptr(), self.transitions.data_ptr(), transition_capacity, self._counters.data_ptr(),
                                    self.reuse_slab.data_ptr() if cfg.reuse_data else None, self.reuse_count.data_ptr() if cfg.reuse_data else None,
                                    self.reuse_capacity, self.ep_stats.data_ptr(),
                                    _C.byref(h)))
        self.h = h
        self.ticks = 0
        self.last_emitted = None         # (slab [n,92], emit_flag [n]) of the latest tick

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
        else:
            self._check(self._L.mopa_rollout_step(self.h, ac.data_ptr(), self._stream()))
            self._keep = ac
        self.ticks += 1
        self.last_emitted = (self.slab, self.emit_flag)
        # relabelled records of this tick: (slab [2n, 92], flags [2n]) - the first reuse_count rows are records
        self.last_reused = (self.reuse_slab, (self._reuse_iota < self.reuse_count).to(self.torch.uint8)) if self.cfg.reuse_data else None

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
    runner.close()
    return info
