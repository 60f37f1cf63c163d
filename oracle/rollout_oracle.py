"""TEST INFRASTRUCTURE — scalar restatement of the reference's experience-collection loop for ONE
environment, built on the CPU oracles (PushEnvOracle, OracleScene, OraclePlanner).  Follows:

  MoPARolloutRunner.run                 rl/mopa_rollouts.py:22-399   (train branch, every_steps=1,
                                                                      no IK; discrete_action and reuse_data on request)
  SACAgent.is_planner_ac / convert2planner_displacement / plan / clip_qpos / simple_interpolate
                                        rl/sac_agent.py:148-318
  PlannerAgent.plan                     rl/planner_agent.py:42-52
  SamplingBasedPlanner.plan             motion_planners/sampling_based_planner.py:60-101

It yields the same SMDP transition records the vectorised runner emits (92 floats) so the two can
be compared env by env, and it is the "reference path on host cores" that bench.py times.
"""
from __future__ import annotations

import numpy as np

from .env_oracle import AssemblyEnvOracle, LiftEnvOracle, PusherEnvOracle, PushEnvOracle
from .oracle import OraclePlanner, OracleScene, space_from_model


class ScalarMoPARunner:
    def __init__(self, model, dynmodel, cfg, ignored, passive, env_gid, seed_env, policy, contacts=True, max_episode_steps=250, task="push"):
        """policy(env_gid, macro_index) -> action (7,) in [-1, 1]."""
        self.m, self.cfg, self.gid, self.policy = model, cfg, int(env_gid), policy
        self.task = task
        env_cls = {"assembly": AssemblyEnvOracle, "lift": LiftEnvOracle, "pusher": PusherEnvOracle}.get(task, PushEnvOracle)
        if task == "pusher":   # BASELINE configs[0]: 4 hinges (joint0 unlimited -> SO(2)), env._ac_scale = 0.1 (pusher_obstacle.py:33)
            self.env = env_cls(model, dynmodel, max_episode_steps=max_episode_steps, contacts=contacts)
        else:
            self.env = env_cls(model, dynmodel, max_episode_steps=max_episode_steps, contacts=contacts, ac_scale=cfg.ac_scale)
        self.scene = OracleScene(model, ignored, cfg.contact_threshold, "f32")
        adr, lo, hi, so2 = space_from_model(model, passive)
        self.planner = OraclePlanner(self.scene, adr, lo, hi, so2, cfg.range, 0.005, cfg.seed, max_nodes=4096)
        # SACAgent._simple_planner (rl/sac_agent.py:98-110): same scene and space, range = simple_planner_range
        self.simple_planner = OraclePlanner(self.scene, adr, lo, hi, so2, cfg.simple_planner_range, 0.005, cfg.seed, max_nodes=4096)
        self.ref = adr
        self.na = len(adr)                                          # arm joints (7 Sawyer, 4 Pusher), qpos addresses 0 .. na-1
        assert list(adr) == list(range(self.na))
        jid = [list(model.jnt_qposadr).index(a) for a in adr]
        self.limited = np.asarray(model.jnt_limited)[jid].astype(bool)
        # unlimited joints (Pusher joint0): never clipped (rl/mopa_rollouts.py:121-131, sac_agent.clip_qpos), wrapped for the planner
        self.jlo = np.where(self.limited, model.jnt_range[jid, 0], -np.inf)
        self.jhi = np.where(self.limited, model.jnt_range[jid, 1], np.inf)
        self.seed_env = seed_env
        self.episode = 0
        self.plan_count = 0
        self.macro_index = 0
        self.env_steps = 0
        self.counters = dict(mp=0, rl=0, interpolation=0, mp_fail=0, approximate=0, invalid=0, reused=0, fb_simple=0, fb_main=0, densify_fallback=0)
        self.contact_force_sum = 0.0   # run_episode: total_contact_force += env.get_contact_force() after every simulated env.step (:538-539, 636-646)
        self.extra_records = []   # relabelled records (reuse_data) of the latest macro step
        if getattr(cfg, "use_ik_target", False):   # MoPA-SAC IK presets: _cart2dispalcement on the oracle's kinematic chain
            from mopa_rl_b200.inverse_kinematics import site_frame   # data-format helper only (site -> simulated body, local position)
            from oracle.ik_oracle import IKOracle

            body, local = site_frame(model, dynmodel, cfg.ik_target)
            dofs = [list(dynmodel.dof_vadr).index(model.get_joint_qvel_addr("right_j%d" % k)) for k in range(7)]
            self.ik = IKOracle(dynmodel, body, local, dofs)
        self.ob = self._reset()

    def _reset(self):
        from mopa_rl_b200.envs import assembly_reset_state, lift_reset_state, push_reset_state  # reset draws are input data shared with the product

        fn = {"assembly": assembly_reset_state, "lift": lift_reset_state}.get(self.task, push_reset_state)
        if self.task == "pusher":
            self.episode += 1
            return self.env.reset(self.seed_env, self.gid, self.episode - 1)
        q, v = fn(self.m, self.seed_env, [self.gid], [self.episode])
        self.episode += 1
        return self.env.reset_to(q[0], v[0])

    def _valid(self, q):
        return bool(self.scene.is_valid(np.asarray(q, np.float64).astype(np.float32).astype(np.float64))[0] & 1)

    def _wrap(self, q):
        """util/env.py:15-25 joint_convert on the unlimited joints (period 3.14, sign preserving); identity for the Sawyer scenes."""
        q = np.array(q, np.float64)
        for k in np.nonzero(~self.limited)[0]:
            period = 3.14 if q[k] > 0 else -3.14
            w = q[k] % period
            if (q[k] // period) % 2 != 0:
                w -= period
            q[k] = w
        return q

    def _clip_qpos(self, q):
        arm = q[:self.na]
        if np.any(arm < self.jlo) or np.any(arm > self.jhi):
            q = q.copy()
            q[:self.na] = np.clip(arm, self.jlo + self.cfg.joint_margin, self.jhi - self.cfg.joint_margin)
        return q

    def _simple_interpolate(self, curr, target):
        """rl/sac_agent.py:262-318 with use_planner=False.  Returns (traj, valid)."""
        lim = self.cfg.ac_scale * 0.8
        curr = self._clip_qpos(curr)
        diff = target[:self.na] - curr[:self.na]
        sf = max(np.max(np.abs(diff) / lim), 1.0)
        scaled = diff / sf
        traj, interp = [], curr.copy()
        for _ in range(int(sf)):
            interp[:self.na] += scaled
            if not self._valid(interp):
                return traj, False
            traj.append(interp.copy())
        traj.append(target)
        return traj, True

    def _rebase(self, curr, states):
        """SamplingBasedPlanner.plan (:72-101) + PlannerAgent.plan (:47-48): the planner's states re-based on the (un-wrapped)
        start, first row dropped."""
        if self.limited.all():
            return [curr + (s - states[0]) for s in states][1:]
        # accumulate waypoint deltas on the un-wrapped start, going the short way round across +-3.14
        path, prev_s, acc = [], states[0], curr.copy()
        for st in states[1:]:
            delta = st - prev_s
            for k in np.nonzero(~self.limited)[0]:
                if abs(st[k] - prev_s[k]) > 3.14:
                    delta[k] = (3.14 - prev_s[k] + st[k] + 3.14) if prev_s[k] > 0 and st[k] <= 0 else (
                        -(3.14 - st[k] + prev_s[k] + 3.14) if prev_s[k] < 0 and st[k] > 0 else delta[k])
            acc = acc + delta
            path.append(acc.copy())
            prev_s = st
        return path

    def _plan(self, curr, target):
        """SACAgent.plan: interpolation first, RRT-Connect when the straight line is blocked, densify."""
        cfg = self.cfg
        curr = self._clip_qpos(curr)
        traj, ok = self._simple_interpolate(curr, target)
        if ok:
            return traj, True, True, True
        key = (self.gid << 32) + self.plan_count
        self.plan_count += 1
        ws, wg = self._wrap(curr), self._wrap(target)               # SamplingBasedPlanner.plan: convert_nonlimited on copies (:63-66)
        r = self.planner.plan(ws.astype(np.float32).astype(np.float64), wg.astype(np.float32).astype(np.float64), key, cfg.max_iter, cfg.max_path)
        if r["status"] != 0:
            return None, False, False, r["status"] != -4
        path = self._rebase(curr, r["path"])
        if cfg.interpolation:
            new, start = [], curr
            for hop, p in enumerate(path):
                d = p[:self.na] - start[:self.na]
                if np.any(np.abs(d) > cfg.ac_scale):
                    lim = cfg.ac_scale * 0.8
                    sf = max(np.max(np.abs(d) / lim), 1.0)
                    inner, interp, good = [], start.copy(), True
                    for _ in range(min(int(sf), int(cfg.range / lim) + 1)):
                        interp[:self.na] += d / sf
                        if not self._valid(interp):
                            good = False
                            break
                        inner.append(interp.copy())
                    mod = int(getattr(cfg, "debug_block_mod", 0))
                    if mod > 0 and min(int(sf), int(cfg.range / lim) + 1) > 0 and (key + hop) % mod == 0:
                        good = False
                    if good:
                        new.extend(inner + [p])
                    else:
                        new.extend(self._replan_hop(start, p, key, hop))
                else:
                    new.append(p)
                start = p
            path = new
        if len(path) > cfg.max_traj:
            return None, False, False, True
        return path, True, False, True

    def _replan_hop(self, start, target, key, hop):
        """simple_interpolate(..., use_planner=True) on a blocked hop (rl/sac_agent.py:300-311): the simple planner
        (simple_planner_timelimit -> cfg.simple_max_iter iterations), then the main planner, else [target]."""
        cfg = self.cfg
        fbkey = ((int(key) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF) ^ (hop + 1)
        s32 = self._wrap(start).astype(np.float32).astype(np.float64)
        g32 = self._wrap(target).astype(np.float32).astype(np.float64)
        r = self.simple_planner.plan(s32, g32, fbkey, cfg.simple_max_iter, 64)
        which = "fb_simple"
        if r["status"] != 0:
            r = self.planner.plan(s32, g32, fbkey ^ 0xA5A5A5A5A5A5A5A5, cfg.max_iter, 64)
            which = "fb_main"
        if r["status"] != 0 or len(r["path"]) < 2:
            self.counters["densify_fallback"] += 1
            return [target]
        self.counters[which] += 1
        return self._rebase(start, r["path"])

    def macro_step(self):
        """One iteration of the `while not done` loop of run(); returns the 92-float transition record."""
        cfg, env = self.cfg, self.env
        if env.terminal:
            self.ob = self._reset()
        prev_ob = self.ob.copy()
        ac = self.policy(self.gid, self.macro_index)
        discrete = bool(getattr(cfg, "discrete_action", False))
        if discrete:                                               # ac = {"default": ..., "ac_type": ...}
            ac, ac_type = ac
        ac = np.asarray(ac, np.float64).astype(np.float32).astype(np.float64)
        lift = self.task == "lift"                                 # 8-D action: 7 joint entries + gripper
        grip_ac = float(ac[7]) if lift else None
        use_ik = bool(getattr(cfg, "use_ik_target", False))
        policy_ac = ac.copy()                                      # what the record keeps
        curr = env.qpos.copy()
        if use_ik:                                                 # ac = (default[3], quat[4], gripper): rl/mopa_rollouts.py:90-101
            ac = np.concatenate([self._cart2displacement(ac, curr), [ac[7]]])
        ac = ac[:self.na]
        self.macro_index += 1
        is_mp = bool(ac_type) if discrete else bool(np.any(np.abs(ac) > cfg.omega))   # rl/mopa_rollouts.py:86-88, 104-111
        self._ac_type = float(is_mp) if discrete else 0.0
        steps = 0
        extra_src = None
        self.extra_records = []
        if is_mp:
            w = cfg.omega
            with np.errstate(divide="ignore", invalid="ignore"):   # omega = 0 (discrete presets): the first branch is never selected
                disp = np.where(np.abs(ac) < w, ac / (w / cfg.ac_scale),
                                np.sign(ac) * (cfg.ac_scale + (cfg.action_range - cfg.ac_scale) * ((np.abs(ac) - w) / (1 - w))))
            if getattr(cfg, "ac_space_type", "piecewise") == "normal":   # rl/sac_agent.py:160-163
                disp = ac * cfg.action_range
            target = curr.copy()
            if not use_ik:                                         # the increment sits behind `if not config.use_ik_target` (:114-131)
                target[:self.na] = np.clip(curr[:self.na] + disp, self.jlo, self.jhi)
            if cfg.invalid_target_handling and not self._valid(target):
                trial = 0
                while not self._valid(target) and trial < cfg.num_trials:
                    d = curr - target
                    target = target + cfg.step_size * d / np.linalg.norm(d)
                    trial += 1
            if self._valid(target):
                traj, success, interpolation, exact = self._plan(curr, target)
                valid = True
            else:
                traj, success, interpolation, valid, exact = None, False, False, False, True
            if success:
                self.counters["interpolation" if interpolation else "mp"] += 1
                meta, done = 0.0, False
                ob_list, rew_list, done_list = [], [], []
                grip_q0 = env.qpos[env.grip_q[0]] if lift else 0.0
                for i, nq in enumerate(traj):
                    a = np.asarray(nq[:self.na] - env.qpos[:self.na], np.float32).astype(np.float64)   # form_action (fp32 action row)
                    if lift:   # form_action's gripper entry (waypoints carry the start state's passive dims), policy's on the last one
                        g = grip_ac if i == len(traj) - 1 else grip_q0 - env.qpos[env.grip_q[0]]
                        a = np.concatenate([a, [np.float64(np.float32(g))]])
                    self.ob, rew, done = env.step(a, is_planner=True)
                    self.contact_force_sum += getattr(env, "contact_force", 0.0)
                    meta += cfg.discount_factor ** i * rew
                    ob_list.append(self.ob.copy()), rew_list.append(meta), done_list.append(done)
                    steps += 1
                    if done:
                        break
                rec_rew, intra = meta, i
                if getattr(cfg, "reuse_data", False) and len(ob_list) > 3:
                    extra_src = (ob_list, rew_list, done_list, traj)
            else:
                self.counters["mp_fail"] += 1
                if not valid:
                    self.counters["invalid"] += 1
                if not exact:
                    self.counters["approximate"] += 1
                rec_rew, done = env.null_step()
                intra, steps = 0, 1
        else:
            self.counters["rl"] += 1
            direct = ac if discrete else (ac / cfg.omega).astype(np.float32).astype(np.float64)   # :347-352
            if lift:
                direct = np.concatenate([direct, [grip_ac]])
            self.ob, rec_rew, done = env.step(direct, is_planner=False)
            self.contact_force_sum += getattr(env, "contact_force", 0.0)
            intra, steps = 0, 1
        env.prev_state = None                                       # env._reset_prev_state()
        self.env_steps += steps
        if extra_src is not None:
            self._reuse(*extra_src)
        rec = np.zeros(92, np.float32)
        no = len(prev_ob)   # 40 (push) / 38 (assembly): observation rows keep the 40-float stride
        rec[0:no], rec[40:40 + self.na], rec[48], rec[49], rec[50], rec[52:52 + no] = prev_ob, ac, rec_rew, float(done), intra, self.ob
        rec[47] = grip_ac if lift else self._ac_type
        if use_ik:
            rec[40:48] = policy_ac
        return rec

    def _cart2displacement(self, ac, curr):
        """MoPARolloutRunner._cart2dispalcement (rl/mopa_rollouts.py:683-728) with util.env.mat2quat (:232-289) spelled out: the
        float32 copy of the site matrix, the eigenvector of K for the largest eigenvalue, w >= 0, returned as (x, y, z, w); then the
        reference's index list [3, 0, 1, 1]; quat_mul = mju_mulQuat; qpos_from_site_pose(max_steps, tol) on the arm joints."""
        cfg = self.cfg
        sp, R, _ = self.ik.site_pose(curr)
        lo, hi = np.array([-1.2, -1.2, 0.0]), np.array([1.2, 1.2, 2.0])          # SawyerEnv.min_world_size / max_world_size (sawyer.py:52-53)
        target_cart = np.clip(sp + cfg.action_range * ac[:3], lo, hi)
        M = np.array(R, dtype=np.float32)
        m00, m01, m02, m10, m11, m12, m20, m21, m22 = [float(x) for x in M.ravel()]
        K = np.array([[m00 - m11 - m22, 0.0, 0.0, 0.0], [m01 + m10, m11 - m00 - m22, 0.0, 0.0],
                      [m02 + m20, m12 + m21, m22 - m00 - m11, 0.0], [m21 - m12, m02 - m20, m10 - m01, m00 + m11 + m22]]) / 3.0
        w, V = np.linalg.eigh(K)
        q = V[[3, 0, 1, 2], np.argmax(w)]
        if q[0] < 0.0:
            q = -q
        q = q[[1, 2, 3, 0]]                                                       # (x, y, z, w)
        target_quat = q[[3, 0, 1, 1]]
        aq = np.asarray(ac[3:7], np.float32)
        aq = (aq / np.linalg.norm(aq)).astype(np.float64)
        from oracle.ik_oracle import _qmul
        target_quat = _qmul(target_quat, aq)
        qres, _, _, _ = self.ik.solve(curr, target_cart, target_quat, max_steps=int(cfg.ik_max_steps), tol=float(cfg.ik_tol))
        tgt = np.clip(qres[:self.na], self.jlo, self.jhi)
        return (tgt - curr[:self.na]).astype(np.float32).astype(np.float64)      # the device front end hands the displacement over as fp32

    def _reuse(self, ob_list, rew_list, done_list, traj):
        """rl/mopa_rollouts.py:223-302: resample (start, goal) waypoint pairs of the executed plan; the reference's
        np.random.randint draws are replaced by the counter-based generator keyed by (env id, macro-action index)."""
        from mopa_rl_b200 import rng

        cfg, L = self.cfg, len(ob_list)
        seed = (int(cfg.seed) + 0x5EED) & 0xFFFFFFFFFFFFFFFF
        mi, pairs = self.macro_index - 1, []
        for t in range(min(L, cfg.max_reuse_data)):
            u0 = float(rng.uniform01(seed, np.uint64(self.gid), np.uint64(mi), np.uint64(2 * t)))
            u1 = float(rng.uniform01(seed, np.uint64(self.gid), np.uint64(mi), np.uint64(2 * t + 1)))
            start = min(int(u0 * (L - 1)), L - 2)
            goal = min(start + 1 + int(u1 * (L - 1 - start)), L - 1)
            if (start, goal) in pairs:
                continue
            pairs.append((start, goal))
            d = traj[goal][:self.na] - traj[start][:self.na]                                  # env.form_action(traj[goal], traj[start])
            s, w, ar = cfg.ac_scale, cfg.omega, cfg.action_range                  # SACAgent.invert_displacement, piecewise
            a = np.where(np.abs(d) < s, d * (w / s), np.sign(d) * ((np.abs(d) - s) / ((ar - s) / (1.0 - s)) / ((1.0 - s) / (1.0 - w)) + w))
            if getattr(cfg, "ac_space_type", "piecewise") == "normal":            # rl/sac_agent.py:180-181
                a = d / ar
            if not (np.any(a < -w) or np.any(a > w)) or not (np.all(a >= -1.0) and np.all(a <= 1.0)):
                continue
            rec = np.zeros(92, np.float32)
            no = len(ob_list[start])
            rec[0:no], rec[40:40 + self.na] = ob_list[start], a
            rec[47] = self._ac_type                                                # inter_subgoal_ac["ac_type"] = ac["ac_type"]
            rec[48] = (rew_list[goal] - rew_list[start]) * cfg.discount_factor ** (-(start + 1))
            rec[49], rec[50], rec[52:52 + no] = float(done_list[goal]), goal - start - 1, ob_list[goal]
            self.extra_records.append(rec)
            self.counters["reused"] += 1
