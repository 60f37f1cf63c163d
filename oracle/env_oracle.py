"""TEST INFRASTRUCTURE — numpy restatement of the SawyerPushObstacle-v0 env logic around the
physics oracle (orc_dyn.c).  Follows, line by line:
  SawyerPushObstacleEnv._step           env/sawyer/sawyer_push_obstacle.py:162-208
  SawyerPushObstacleEnv.compute_reward  :71-100
  SawyerEnv._get_obs + push _get_obs    env/sawyer/sawyer.py:317-338, sawyer_push_obstacle.py:102-116
  BaseEnv.step / _after_step            env/base.py:232-314
One instance = one environment, exactly like the reference.  Body / site frames read after a step
are those of the last mj_step's START state (mjData is not refreshed after integration).
"""
from __future__ import annotations

import numpy as np

from .oracle import OracleDyn


def _q2m(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


class PushEnvOracle:
    def __init__(self, model, dynmodel, max_episode_steps=250, frame_dt=0.15, ac_scale=0.05, distance_threshold=0.06,
                 success_reward=150.0, contacts=True):
        self.m, self.dm = model, dynmodel
        self.dyn = OracleDyn(dynmodel)
        self.dyn.enable_contacts(contacts)
        m = model
        self.ref_q = [m.get_joint_qpos_addr("right_j%d" % i) for i in range(7)]
        self.ref_v = [m.get_joint_qvel_addr("right_j%d" % i) for i in range(7)]
        self.grip_q = [m.get_joint_qpos_addr(j) for j in ("rc_close", "lc_close")]
        self.grip_v = [m.get_joint_qvel_addr(j) for j in ("rc_close", "lc_close")]
        self.tgt_q = [m.get_joint_qpos_addr(j) for j in ("target_x", "target_y")]
        sim = {b: i for i, b in enumerate(dynmodel.bodies)}
        self.b_ee, self.b_cube = sim[m.body_name2id("right_ee_attchment")], sim[m.body_name2id("cube")]
        self.b_rc, self.b_lc = sim[m.body_name2id("rightclaw")], sim[m.body_name2id("leftclaw")]
        self.s_re, self.s_le, self.s_grip = (m.site_pos[m.site_name2id(n)] for n in ("right_eef", "left_eef", "grip_site"))
        self.target_base = m.body_pos[m.body_name2id("target")]
        self.comp = np.zeros(dynmodel.nd, np.int32)
        self.comp[[list(dynmodel.dof_vadr).index(v) for v in self.ref_v]] = 1
        self.nsub = int(frame_dt / m.opt_timestep)
        self.ac_scale, self.dthr, self.succ_rew, self.max_steps = ac_scale, distance_threshold, success_reward, max_episode_steps
        self.lim = [(int(q), dynmodel._arr["d_range"][k]) for k, q in enumerate(dynmodel.dof_qadr) if q >= 0 and dynmodel._arr["d_limited"][k]]

    def set_state(self, qpos, qvel):
        self.qpos, self.qvel = np.array(qpos, np.float64), np.array(qvel, np.float64)
        self.bias_prev, self.xpos, self.xquat = self.dyn.forward(self.qpos, self.qvel)

    def reset_to(self, qpos, qvel):
        self.set_state(qpos, qvel)
        self.prev_state = None
        self.ep_len, self.ep_rew, self.terminal, self.success = 0, 0.0, False, False
        return self.obs()

    def _site(self, b, local):
        return self.xpos[b] + _q2m(self.xquat[b]) @ local

    def obs(self):
        q, v = self.qpos, self.qvel
        eef = self._site(self.b_ee, self.s_grip)
        target = self.target_base + np.array([q[self.tgt_q[0]], q[self.tgt_q[1]], 0.0])
        cube, cq, eq = self.xpos[self.b_cube], self.xquat[self.b_cube], self.xquat[self.b_ee]
        return np.concatenate([q[self.ref_q], v[self.ref_v], q[self.grip_q], v[self.grip_v], eef, eq[[1, 2, 3, 0]], target, cube,
                               cq[[1, 2, 3, 0]], eef - cube, cube[:2] - target[:2]])

    def _reward(self):
        """SawyerPushObstacleEnv.compute_reward (:71-100) -> (reward, terminal)."""
        gs = 0.5 * (self._site(self.b_rc, self.s_re) + self._site(self.b_lc, self.s_le))
        cube = self.xpos[self.b_cube]
        target = self.target_base + np.array([self.qpos[self.tgt_q[0]], self.qpos[self.tgt_q[1]], 0.0])
        d_gc, d_ct = np.linalg.norm(cube - gs), np.linalg.norm(cube[:2] - target[:2])
        reward = 0.0
        if d_ct < 0.1:
            reward += 0.5 * (1 - np.tanh(5 * d_ct))
        if d_gc < 0.1:
            reward += 0.1 * (1 - np.tanh(10 * d_gc))
        terminal = False
        if d_ct < self.dthr:
            reward += self.succ_rew
            self.success, terminal = True, True
        return reward, terminal

    def step(self, action, is_planner=False):
        action = np.asarray(action, np.float64)
        if not is_planner or self.prev_state is None:
            self.prev_state = self.qpos[self.ref_q].copy()
        a = action[:7] if is_planner else action[:7] * self.ac_scale
        desired = self.prev_state + np.clip(a, -self.ac_scale, self.ac_scale)
        self.qpos, self.qvel, self.bias_prev, self.xpos, self.xquat, self.ncon = self.dyn.step(
            self.qpos, self.qvel, desired, self.comp, self.bias_prev, self.nsub)
        self.contact_force = self.dyn.contact_force   # BaseEnv.get_contact_force()
        self.prev_state = desired.copy()
        reward, terminal = self._reward()   # compute_reward
        ob = self.obs()
        # _after_step
        clipped = False
        for qa, (lo, hi) in self.lim:
            if self.qpos[qa] < lo or self.qpos[qa] > hi:
                self.qpos[qa] = min(max(self.qpos[qa], lo), hi)
                clipped = True
        if clipped:
            self.set_state(self.qpos, self.qvel)
        self.ep_rew += reward
        self.ep_len += 1
        if self.ep_len == self.max_steps:
            terminal = True
        self.terminal = terminal
        return ob, reward, terminal

    def null_step(self):
        """Planner failure (rl/mopa_rollouts.py:304-327): compute_reward(zeros) + _after_step, no simulation.
        Frames are refreshed by a forward pass first (the device path does the same)."""
        self.set_state(self.qpos, self.qvel)
        reward, terminal = self._reward()
        self.ep_rew += reward
        self.ep_len += 1
        if self.ep_len == self.max_steps:
            terminal = True
        self.terminal = terminal
        return reward, terminal


class AssemblyEnvOracle(PushEnvOracle):
    """SawyerAssemblyObstacle-v0 (env/sawyer/sawyer_assembly_obstacle.py:32-59, _step :97-143): same step and
    _after_step logic as the push task; reward from the pegHead / hole / hole_bottom sites, 38-float observation."""

    def __init__(self, model, dynmodel, max_episode_steps=250, frame_dt=0.15, ac_scale=0.05, success_reward=150.0, contacts=True):
        self.m, self.dm = model, dynmodel
        self.dyn = OracleDyn(dynmodel)
        self.dyn.enable_contacts(contacts)
        m = model
        self.ref_q = [m.get_joint_qpos_addr("right_j%d" % i) for i in range(7)]
        self.ref_v = [m.get_joint_qvel_addr("right_j%d" % i) for i in range(7)]
        self.grip_q = [m.get_joint_qpos_addr(j) for j in ("rc_close", "lc_close")]
        self.grip_v = [m.get_joint_qvel_addr(j) for j in ("rc_close", "lc_close")]
        sim = {b: i for i, b in enumerate(dynmodel.bodies)}
        self.b_ee, self.b_peg = sim[m.body_name2id("right_ee_attchment")], sim[m.body_name2id("peg")]
        self.b_hole = sim[int(m.site_bodyid[m.site_name2id("hole")])]
        self.s_head, self.s_end, self.s_grip, self.s_hole, self.s_bottom = (
            m.site_pos[m.site_name2id(n)] for n in ("pegHead", "pegEnd", "grip_site", "hole", "hole_bottom"))
        self.comp = np.zeros(dynmodel.nd, np.int32)
        self.comp[[list(dynmodel.dof_vadr).index(v) for v in self.ref_v]] = 1
        self.nsub = int(frame_dt / m.opt_timestep)
        self.ac_scale, self.succ_rew, self.max_steps = ac_scale, success_reward, max_episode_steps
        self.lim = [(int(q), dynmodel._arr["d_range"][k]) for k, q in enumerate(dynmodel.dof_qadr) if q >= 0 and dynmodel._arr["d_limited"][k]]

    def obs(self):
        q, v = self.qpos, self.qvel
        eef, eq = self._site(self.b_ee, self.s_grip), self.xquat[self.b_ee]
        return np.concatenate([q[self.ref_q], v[self.ref_v], q[self.grip_q], v[self.grip_v], eef, eq[[1, 2, 3, 0]],
                               self._site(self.b_hole, self.s_hole), self._site(self.b_peg, self.s_head), self._site(self.b_peg, self.s_end),
                               self.xquat[self.b_peg]])

    def _reward(self):
        head = self._site(self.b_peg, self.s_head)
        d_hole = np.linalg.norm(head - self._site(self.b_hole, self.s_hole))
        d_bottom = np.linalg.norm(head - self._site(self.b_hole, self.s_bottom))
        reward, terminal = 0.0, False
        if d_hole < 0.3:
            reward += 0.4 * (1 - np.tanh(15 * d_hole))
        if d_bottom < 0.025:
            reward += self.succ_rew
            self.success, terminal = True, True
        return reward, terminal


class LiftEnvOracle(PushEnvOracle):
    """SawyerLiftObstacle-v0 (env/sawyer/sawyer_lift_obstacle.py): 8-D action (7 joints + gripper, _step :191-240 with
    SawyerEnv._gripper_format_action :340-342), reward = max(reach, grasp, lift) with has_grasp read from the contact
    list of the last mj_step (:92-148), 35-float observation (:150-161)."""
    LEFT = ("l_finger_g0", "l_finger_g1", "l_fingertip_g0")
    RIGHT = ("r_finger_g0", "r_finger_g1", "r_fingertip_g0")

    def __init__(self, model, dynmodel, max_episode_steps=250, frame_dt=0.15, ac_scale=0.05, success_reward=150.0, contacts=True):
        self.m, self.dm = model, dynmodel
        self.dyn = OracleDyn(dynmodel)
        self.dyn.enable_contacts(contacts)
        m = model
        self.ref_q = [m.get_joint_qpos_addr("right_j%d" % i) for i in range(7)]
        self.ref_v = [m.get_joint_qvel_addr("right_j%d" % i) for i in range(7)]
        self.grip_q = [m.get_joint_qpos_addr(j) for j in ("rc_close", "lc_close")]
        self.grip_v = [m.get_joint_qvel_addr(j) for j in ("rc_close", "lc_close")]
        sim = {b: i for i, b in enumerate(dynmodel.bodies)}
        self.b_ee, self.b_cube = sim[m.body_name2id("right_ee_attchment")], sim[m.body_name2id("cube")]
        self.s_grip = m.site_pos[m.site_name2id("grip_site")]
        self.g_cube = m.geom_name2id("cube")
        self.g_left, self.g_right = [m.geom_name2id(n) for n in self.LEFT], [m.geom_name2id(n) for n in self.RIGHT]
        self.bin_z = float(m.body_pos[m.body_name2id("bin1")][2])
        self.comp = np.zeros(dynmodel.nd, np.int32)
        self.comp[[list(dynmodel.dof_vadr).index(v) for v in self.ref_v]] = 1
        self.nsub = int(frame_dt / m.opt_timestep)
        self.ac_scale, self.succ_rew, self.max_steps = ac_scale, success_reward, max_episode_steps
        self.lim = [(int(q), dynmodel._arr["d_range"][k]) for k, q in enumerate(dynmodel.dof_qadr) if q >= 0 and dynmodel._arr["d_limited"][k]]
        self.contacts = []

    def obs(self):
        q, v = self.qpos, self.qvel
        eef, eq = self._site(self.b_ee, self.s_grip), self.xquat[self.b_ee]
        cube, cq = self.xpos[self.b_cube], self.xquat[self.b_cube]
        return np.concatenate([q[self.ref_q], v[self.ref_v], q[self.grip_q], v[self.grip_v], eef, eq[[1, 2, 3, 0]], cube, cq[[1, 2, 3, 0]],
                               eef - cube])

    def _reward(self):
        grip, cube = self._site(self.b_ee, self.s_grip), self.xpos[self.b_cube]
        reach = (1 - np.tanh(10 * np.linalg.norm(cube - grip))) * 0.1
        tl = tr = False
        for g1, g2 in self.contacts:
            other = g2 if g1 == self.g_cube else (g1 if g2 == self.g_cube else None)
            tl, tr = tl or other in self.g_left, tr or other in self.g_right
        self.has_grasp = tl and tr
        grasp = 0.35 if self.has_grasp else 0.0
        z_target = self.bin_z + 0.45
        lift = 0.35 + (1 - np.tanh(15 * max(z_target - cube[2], 0.0))) * (0.5 - 0.35) if self.has_grasp else 0.0
        reward, terminal = max(reach, grasp, lift), False
        if self.has_grasp and abs(cube[2] - z_target) < 0.05:
            reward += self.succ_rew
            self.success, terminal = True, True
        return reward, terminal

    def step(self, action, is_planner=False):
        action = np.asarray(action, np.float64)
        assert len(action) == 8
        if not is_planner or self.prev_state is None:
            self.prev_state = self.qpos[self.ref_q].copy()
        a = action[:7] if is_planner else action[:7] * self.ac_scale
        desired = self.prev_state + np.clip(a, -self.ac_scale, self.ac_scale)
        ctrl = np.concatenate([desired, self.qpos[self.grip_q] + action[7]])
        self.qpos, self.qvel, self.bias_prev, self.xpos, self.xquat, self.ncon = self.dyn.step(
            self.qpos, self.qvel, ctrl, self.comp, self.bias_prev, self.nsub)
        self.contact_force, self.contacts = self.dyn.contact_force, self.dyn.contacts
        self.prev_state = desired.copy()
        reward, terminal = self._reward()
        ob = self.obs()
        clipped = False
        for qa, (lo, hi) in self.lim:
            if self.qpos[qa] < lo or self.qpos[qa] > hi:
                self.qpos[qa] = min(max(self.qpos[qa], lo), hi)
                clipped = True
        if clipped:
            self.set_state(self.qpos, self.qvel)
        self.ep_rew += reward
        self.ep_len += 1
        if self.ep_len == self.max_steps:
            terminal = True
        self.terminal = terminal
        return ob, reward, terminal

    def reset_to(self, qpos, qvel):
        self.contacts = []   # sim.forward() of the reset state: the can rests in the bin, no finger touches it
        return super().reset_to(qpos, qvel)

    def null_step(self):
        # no mj_step ran: compute_reward(zeros) reads the contact list the previous step left in mjData (rl/mopa_rollouts.py:312)
        return super().null_step()


def pusher_reset_state(model, seed, env_id, episode):
    """PusherObstacleEnv._reset (env/pusher/pusher_obstacle.py:40-67) with counter-based draws keyed by (seed, env id,
    episode, attempt): goal / box ~ U([-0.35, 0.13], [-0.24, 0.2]), qpos0 + U(+-0.02), qvel ~ U(+-0.005) with the goal /
    box velocities zero.  The acceptance test (no contact, box farther than 0.1 from the target, goal[0] <= box[0]) needs
    the collision oracle and is applied by the caller (PusherEnvOracle.reset)."""
    from mopa_rl_b200 import rng

    def draw(attempt):
        s = np.uint64(env_id) * np.uint64(1000003) + np.uint64(attempt)
        u = rng.uniform01(seed, s, np.uint64(episode), np.arange(4 + model.nq + model.nv, dtype=np.uint64))
        lo, hi = np.array([-0.35, 0.13]), np.array([-0.24, 0.2])
        goal, box = lo + (hi - lo) * u[0:2], lo + (hi - lo) * u[2:4]
        qpos = model.qpos0 + (-0.02 + 0.04 * u[4:4 + model.nq])
        qpos[-4:-2], qpos[-2:] = goal, box
        qvel = -0.005 + 0.01 * u[4 + model.nq:]
        qvel[-4:] = 0.0
        return qpos, qvel, goal, box

    return draw


class PusherEnvOracle:
    """PusherObstacle-v0 (BASELINE configs[0], the reference's CPU-runnable case): 4 hinge joints driven through velocity
    actuators (gear 10) by the PID law of BaseEnv._get_control (env/base.py:200-209: kp 150, kd 20, ki 0.1, leak 0.95),
    RK4 with dt = 0.01, int(frame_dt / dt) = 100 mj_steps per env.step (env/pusher/pusher_obstacle.py:240-284), reward
    :223-238, observation :185-205, BaseEnv._after_step.  `desired_state = prev_state + action` - the clipped / scaled
    variants computed before it are dead code in the reference and are not applied here either."""

    def __init__(self, model, dynmodel, max_episode_steps=150, frame_dt=1.0, distance_threshold=0.05, success_reward=150.0,
                 kp=150.0, kd=20.0, ki=0.1, contacts=True):
        self.m, self.dm = model, dynmodel
        self.dyn = OracleDyn(dynmodel)
        self.dyn.enable_contacts(contacts)
        m = model
        self.ref_q = [m.get_joint_qpos_addr("joint%d" % i) for i in range(4)]
        self.ref_v = [m.get_joint_qvel_addr("joint%d" % i) for i in range(4)]
        sim = {b: i for i, b in enumerate(dynmodel.bodies) if b >= 0}
        self.b_tip, self.b_box, self.b_target = (sim[m.body_name2id(n)] for n in ("fingertip", "box", "target"))
        self.s_tip = m.site_pos[m.site_name2id("fingertip")]
        self.nsub = int(frame_dt / m.opt_timestep)
        self.kp, self.kd, self.ki = kp, kd, ki
        self.dthr, self.succ_rew, self.max_steps, self.ac_scale = distance_threshold, success_reward, max_episode_steps, 0.1
        self.lim = [(int(q), dynmodel._arr["d_range"][k]) for k, q in enumerate(dynmodel.dof_qadr) if q >= 0 and dynmodel._arr["d_limited"][k]]
        self.zero_comp = np.zeros(dynmodel.nd, np.int32)

    def set_state(self, qpos, qvel):
        self.qpos, self.qvel = np.array(qpos, np.float64), np.array(qvel, np.float64)
        _, self.xpos, self.xquat = self.dyn.forward(self.qpos, self.qvel)

    def ncon_at(self, qpos):
        """sim.data.ncon after set_state: contacts at this configuration (one zero-length look through the physics oracle)."""
        q, v = np.array(qpos, np.float64), np.zeros(self.m.nv)
        *_, ncon = self.dyn.step(q, v, np.zeros(4), self.zero_comp, np.zeros(self.dm.nd), 1)
        return ncon

    def reset(self, seed, env_id, episode):
        draw = pusher_reset_state(self.m, seed, env_id, episode)
        for attempt in range(1000):
            qpos, qvel, goal, box = draw(attempt)
            self.set_state(qpos, qvel)
            d = np.linalg.norm(self.xpos[self.b_box] - self.xpos[self.b_target])
            if self.ncon_at(qpos) == 0 and d > 0.1 and goal[0] <= box[0]:
                break
        else:
            raise RuntimeError("no admissible reset state in 1000 draws")
        return self.reset_to(qpos, qvel)

    def reset_to(self, qpos, qvel):
        self.set_state(qpos, qvel)
        self.prev_state, self.i_term = None, np.zeros(4)
        self.ep_len, self.ep_rew, self.terminal, self.success = 0, 0.0, False, False
        return self.obs()

    def obs(self):
        th = self.qpos[self.ref_q]
        tip = self.xpos[self.b_tip] + _q2m(self.xquat[self.b_tip]) @ self.s_tip
        return np.concatenate([np.cos(th), np.sin(th), self.qpos[-2:], self.qvel[self.ref_v], self.qvel[-2:], tip[:2], self.qpos[-4:-2]])

    def _reward(self):
        tip = self.xpos[self.b_tip] + _q2m(self.xquat[self.b_tip]) @ self.s_tip
        d_bg = np.linalg.norm(self.xpos[self.b_box] - tip)
        d_bt = np.linalg.norm(self.xpos[self.b_box] - self.xpos[self.b_target])
        reward = 0.0
        if d_bg < 0.1:
            reward += 0.1 * (1 - np.tanh(5 * d_bg))
        if d_bt < 0.1:
            reward += 0.3 * (1 - np.tanh(5 * d_bt))
        terminal = False
        if d_bt < self.dthr:
            self.success, terminal = True, True
            reward += self.succ_rew
        return reward, terminal

    def step(self, action, is_planner=False):
        action = np.asarray(action, np.float64)
        if not is_planner or self.prev_state is None:
            self.prev_state = self.qpos[self.ref_q].copy()
        desired = self.prev_state + action
        bias = np.zeros(self.dm.nd)
        for _ in range(self.nsub):   # the PID law is re-evaluated before every mj_step
            p = self.kp * (desired - self.qpos[self.ref_q])
            d = self.kd * (0.0 - self.qvel[self.ref_v])
            self.i_term = 0.95 * self.i_term + self.ki * (self.prev_state - self.qpos[self.ref_q])
            self.qpos, self.qvel, bias, self.xpos, self.xquat, self.ncon = self.dyn.step(
                self.qpos, self.qvel, p + d + self.i_term, self.zero_comp, bias, 1)
        self.prev_state = desired.copy()
        reward, terminal = self._reward()
        ob = self.obs()
        clipped = False
        for qa, (lo, hi) in self.lim:
            if self.qpos[qa] < lo or self.qpos[qa] > hi:
                self.qpos[qa] = min(max(self.qpos[qa], lo), hi)
                clipped = True
        if clipped:
            self.set_state(self.qpos, self.qvel)
        self.ep_rew += reward
        self.ep_len += 1
        if self.ep_len == self.max_steps:
            terminal = True
        self.terminal = terminal
        return ob, reward, terminal

    def null_step(self):
        """Planner failure (rl/mopa_rollouts.py:304-327): compute_reward + _after_step, no simulation; frames refreshed first."""
        self.set_state(self.qpos, self.qvel)
        reward, terminal = self._reward()
        self.ep_rew += reward
        self.ep_len += 1
        if self.ep_len == self.max_steps:
            terminal = True
        self.terminal = terminal
        return reward, terminal
