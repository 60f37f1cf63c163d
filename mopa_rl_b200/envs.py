"""Vectorised Sawyer environments on the GPU behind the reference's env semantics.

``VecSawyerPushObstacle`` holds N independent copies of SawyerPushObstacle-v0
(env/sawyer/sawyer_push_obstacle.py) as device arrays (torch tensors) and advances them with
``mopa_env_step`` (one thread per env, 75 substeps on chip).  Reset follows
``SawyerPushObstacleEnv._reset`` (:36-51): arm = init_qpos + N(0, 0.02^2), target slide joints +=
U(-0.01, 0.01), with draws from the counter-based generator keyed by (seed, env id, episode #).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import rng
from .capi import check, lib
from .dynmodel import DynModel
from .model import load_model

PUSH_INIT_QPOS = np.array([0.000457, -0.114, 0.0321, -0.00712, 0.0303, -0.0302, -0.00994])
OBS_KEYS = (("joint_pos", 7), ("joint_vel", 7), ("gripper_qpos", 2), ("gripper_qvel", 2), ("eef_pos", 3), ("eef_quat", 4),
            ("target_pos", 3), ("cube_pos", 3), ("cube_quat", 4), ("gripper_to_cube", 3), ("cube_to_target", 2))
OBS_DIM = 40
BIAS_PAD = 16


class SawyerTask(C.Structure):
    _fields_ = [("kind", C.c_int32), ("arm_qadr", C.c_int32 * 7), ("arm_vadr", C.c_int32 * 7), ("arm_dof", C.c_int32 * 7),
                ("grip_qadr", C.c_int32 * 2), ("grip_vadr", C.c_int32 * 2), ("body_ee", C.c_int32), ("body_cube", C.c_int32),
                ("body_rclaw", C.c_int32), ("body_lclaw", C.c_int32), ("target_qadr", C.c_int32 * 2),
                ("max_episode_steps", C.c_int32), ("nsub", C.c_int32), ("site_right_eef", C.c_double * 3),
                ("site_left_eef", C.c_double * 3), ("site_grip", C.c_double * 3), ("target_base", C.c_double * 3),
                ("ac_scale", C.c_double), ("distance_threshold", C.c_double), ("success_reward", C.c_double),
                ("site_hole", C.c_double * 3), ("site_hole_bottom", C.c_double * 3),
                ("geom_cube", C.c_int32), ("geom_lfinger", C.c_int32 * 3), ("geom_rfinger", C.c_int32 * 3), ("n_arm", C.c_int32),
                ("bin_z", C.c_double), ("unstable_penalty", C.c_double), ("pid_kp", C.c_double), ("pid_kd", C.c_double), ("pid_ki", C.c_double)]


class EnvBuffers(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("qpos", "qvel", "prev_state", "bias_prev", "has_prev", "ep_len", "ep_rew", "obs",
                                           "reward", "done", "success", "ncon", "work", "cforce", "grasp", "i_term", "unstable")]


def make_push_task(model, dyn, max_episode_steps=250, frame_dt=0.15, ac_scale=0.05, distance_threshold=0.06, success_reward=150.0,
                   with_target=True, unstable_penalty=0.0):
    t = SawyerTask()
    t.unstable_penalty = float(unstable_penalty)
    t.kind = 0
    joints = ["right_j%d" % i for i in range(7)]
    sim_body = {b: i for i, b in enumerate(dyn.bodies)}
    for k, j in enumerate(joints):
        t.arm_qadr[k] = model.get_joint_qpos_addr(j)
        t.arm_vadr[k] = model.get_joint_qvel_addr(j)
        t.arm_dof[k] = list(dyn.dof_vadr).index(t.arm_vadr[k])
    for k, j in enumerate(["rc_close", "lc_close"]):
        t.grip_qadr[k] = model.get_joint_qpos_addr(j)
        t.grip_vadr[k] = model.get_joint_qvel_addr(j)
    t.body_ee = sim_body[model.body_name2id("right_ee_attchment")]
    t.body_cube = sim_body[model.body_name2id("cube")]
    t.body_rclaw = sim_body[model.body_name2id("rightclaw")]
    t.body_lclaw = sim_body[model.body_name2id("leftclaw")]
    for k, j in enumerate(["target_x", "target_y"] if with_target else []):
        t.target_qadr[k] = model.get_joint_qpos_addr(j)
    t.max_episode_steps = int(max_episode_steps)
    t.nsub = int(frame_dt / model.opt_timestep)
    for name, field in ((("right_eef", t.site_right_eef), ("left_eef", t.site_left_eef)) if with_target else ()) + (("grip_site", t.site_grip),):
        p = model.site_pos[model.site_name2id(name)]
        for k in range(3):
            field[k] = float(p[k])
    tb = model.body_pos[model.body_name2id("target")] if with_target else np.zeros(3)
    for k in range(3):
        t.target_base[k] = float(tb[k])
    t.geom_cube = -1
    for k in range(3):
        t.geom_lfinger[k] = t.geom_rfinger[k] = -1
    t.ac_scale, t.distance_threshold, t.success_reward = float(ac_scale), float(distance_threshold), float(success_reward)
    return t


ASSEMBLY_INIT_QPOS = np.array([0.427, 0.13, 0.0557, 0.114, -0.0622, 0.0276, 0.00356])   # sawyer_assembly_obstacle.py:19-20


def make_assembly_task(model, dyn, max_episode_steps=250, frame_dt=0.15, ac_scale=0.05, success_reward=150.0, unstable_penalty=0.0, **_):
    """SawyerAssemblyObstacle-v0: peg rigidly attached to the gripper, hole sites on the furniture part `4_part4`
    (env/sawyer/sawyer_assembly_obstacle.py).  kind 2 of mopa_sawyer_task: body_cube = peg, body_rclaw / body_lclaw =
    the body that carries the hole sites, site_right_eef / site_left_eef = pegHead / pegEnd."""
    t = SawyerTask()
    t.unstable_penalty = float(unstable_penalty)
    t.kind = 2
    sim_body = {b: i for i, b in enumerate(dyn.bodies)}
    for k in range(7):
        j = "right_j%d" % k
        t.arm_qadr[k] = model.get_joint_qpos_addr(j)
        t.arm_vadr[k] = model.get_joint_qvel_addr(j)
        t.arm_dof[k] = list(dyn.dof_vadr).index(t.arm_vadr[k])
    for k, j in enumerate(["rc_close", "lc_close"]):
        t.grip_qadr[k] = model.get_joint_qpos_addr(j)
        t.grip_vadr[k] = model.get_joint_qvel_addr(j)
    hole_body = int(model.site_bodyid[model.site_name2id("hole")])
    assert hole_body == int(model.site_bodyid[model.site_name2id("hole_bottom")])
    t.body_ee = sim_body[model.body_name2id("right_ee_attchment")]
    t.body_cube = sim_body[model.body_name2id("peg")]
    t.body_rclaw = t.body_lclaw = sim_body[hole_body]
    t.max_episode_steps = int(max_episode_steps)
    t.nsub = int(frame_dt / model.opt_timestep)
    for name, field in (("pegHead", t.site_right_eef), ("pegEnd", t.site_left_eef), ("grip_site", t.site_grip),
                        ("hole", t.site_hole), ("hole_bottom", t.site_hole_bottom)):
        p = model.site_pos[model.site_name2id(name)]
        for k in range(3):
            field[k] = float(p[k])
    t.ac_scale, t.distance_threshold, t.success_reward = float(ac_scale), 0.0, float(success_reward)
    return t


LIFT_INIT_QPOS = np.array([-0.0305, -0.7325, 0.03043, 1.16124, 1.87488, 0, 0])   # sawyer_lift_obstacle.py:13-14
LIFT_LEFT_FINGER_GEOMS = ("l_finger_g0", "l_finger_g1", "l_fingertip_g0")      # :43-49
LIFT_RIGHT_FINGER_GEOMS = ("r_finger_g0", "r_finger_g1", "r_fingertip_g0")


def make_lift_task(model, dyn, max_episode_steps=250, frame_dt=0.15, ac_scale=0.05, success_reward=150.0, unstable_penalty=0.0, **_):
    """SawyerLiftObstacle-v0 (env/sawyer/sawyer_lift_obstacle.py): kind 1 of mopa_sawyer_task.  8-D action (7 joints +
    gripper), reward = max(reach, grasp, lift) with has_grasp read from the contact list (both fingers touch the can)."""
    t = make_push_task(model, dyn, max_episode_steps, frame_dt, ac_scale, 0.0, success_reward, with_target=False, unstable_penalty=unstable_penalty)
    t.kind = 1
    sim_geom = {g: i for i, g in enumerate(dyn.geoms)}
    t.geom_cube = sim_geom[model.geom_name2id("cube")]
    for k in range(3):
        t.geom_lfinger[k] = sim_geom.get(model.geom_name2id(LIFT_LEFT_FINGER_GEOMS[k]), -1)
        t.geom_rfinger[k] = sim_geom.get(model.geom_name2id(LIFT_RIGHT_FINGER_GEOMS[k]), -1)
    t.bin_z = float(model.body_pos[model.body_name2id("bin1")][2])
    return t


def make_pusher_task(model, dyn, max_episode_steps=400, frame_dt=1.0, distance_threshold=0.05, success_reward=150.0, unstable_penalty=0.0,
                     kp=150.0, kd=20.0, ki=0.1, **_):
    """PusherObstacle-v0 (env/pusher/pusher_obstacle.py; BASELINE configs[0]): kind 3 of mopa_sawyer_task.  4 hinge joints under velocity
    actuators (gear 10) driven by the PID law of BaseEnv._get_control (env/base.py:200-209), RK4 with dt 0.01, int(frame_dt / dt) = 100
    mj_steps per env.step; max_episode_steps 400 as in scripts/2d/mopa.sh."""
    t = SawyerTask()
    t.kind, t.n_arm = 3, 4
    t.unstable_penalty = float(unstable_penalty)
    sim_body = {b: i for i, b in enumerate(dyn.bodies) if b >= 0}
    for k in range(4):
        j = "joint%d" % k
        t.arm_qadr[k] = model.get_joint_qpos_addr(j)
        t.arm_vadr[k] = model.get_joint_qvel_addr(j)
        t.arm_dof[k] = list(dyn.dof_vadr).index(t.arm_vadr[k])
    for k, j in enumerate(["box_x", "box_y"]):          # qpos[-2:] / qvel[-2:]
        t.grip_qadr[k] = model.get_joint_qpos_addr(j)
        t.grip_vadr[k] = model.get_joint_qvel_addr(j)
    for k, j in enumerate(["target_x", "target_y"]):    # qpos[-4:-2]: the goal
        t.target_qadr[k] = model.get_joint_qpos_addr(j)
    assert [t.target_qadr[0], t.target_qadr[1], t.grip_qadr[0], t.grip_qadr[1]] == list(range(model.nq - 4, model.nq))
    t.body_ee = sim_body[model.body_name2id("fingertip")]
    t.body_cube = sim_body[model.body_name2id("box")]
    t.body_rclaw = t.body_lclaw = sim_body[model.body_name2id("target")]
    t.max_episode_steps = int(max_episode_steps)
    t.nsub = int(frame_dt / model.opt_timestep)
    p = model.site_pos[model.site_name2id("fingertip")]
    for k in range(3):
        t.site_grip[k] = float(p[k])
    t.geom_cube = -1
    for k in range(3):
        t.geom_lfinger[k] = t.geom_rfinger[k] = -1
    t.ac_scale, t.distance_threshold, t.success_reward = 0.1, float(distance_threshold), float(success_reward)   # _ac_scale: pusher_obstacle.py:33
    t.pid_kp, t.pid_kd, t.pid_ki = float(kp), float(kd), float(ki)
    return t


def lift_reset_state(model, seed, env_ids, episode_idx):
    """Reset distribution of SawyerLiftObstacleEnv._reset (:23-32): arm = init_qpos + N(0, 0.02^2), the can at its keyframe pose."""
    env_ids = np.asarray(env_ids, dtype=np.uint64).reshape(-1)
    ep = np.broadcast_to(np.asarray(episode_idx, dtype=np.uint64), env_ids.shape)
    n = len(env_ids)
    qpos = np.tile(model.qpos0, (n, 1))
    ref = [model.get_joint_qpos_addr("right_j%d" % i) for i in range(7)]
    dims = np.arange(7, dtype=np.uint64)
    qpos[:, ref] = LIFT_INIT_QPOS + 0.02 * rng.normal(seed, env_ids[:, None], ep[:, None], dims[None, :])
    return qpos, np.zeros((n, model.nv))


def assembly_reset_state(model, seed, env_ids, episode_idx):
    """Reset distribution of SawyerAssemblyObstacleEnv._reset (:22-30): arm = init_qpos + N(0, 0.02^2), rest at qpos0."""
    env_ids = np.asarray(env_ids, dtype=np.uint64).reshape(-1)
    ep = np.broadcast_to(np.asarray(episode_idx, dtype=np.uint64), env_ids.shape)
    n = len(env_ids)
    qpos = np.tile(model.qpos0, (n, 1))
    ref = [model.get_joint_qpos_addr("right_j%d" % i) for i in range(7)]
    dims = np.arange(7, dtype=np.uint64)
    qpos[:, ref] = ASSEMBLY_INIT_QPOS + 0.02 * rng.normal(seed, env_ids[:, None], ep[:, None], dims[None, :])
    return qpos, np.zeros((n, model.nv))


def push_reset_state(model, seed, env_ids, episode_idx):
    """Reset distribution of SawyerPushObstacleEnv._reset for the given (env id, episode #) pairs."""
    env_ids = np.asarray(env_ids, dtype=np.uint64).reshape(-1)
    ep = np.broadcast_to(np.asarray(episode_idx, dtype=np.uint64), env_ids.shape)
    n = len(env_ids)
    qpos = np.tile(model.qpos0, (n, 1))
    ref = [model.get_joint_qpos_addr("right_j%d" % i) for i in range(7)]
    tgt = [model.get_joint_qpos_addr(j) for j in ("target_x", "target_y")]
    dims = np.arange(7, dtype=np.uint64)
    qpos[:, ref] = PUSH_INIT_QPOS + 0.02 * rng.normal(seed, env_ids[:, None], ep[:, None], dims[None, :])
    u = rng.uniform01(seed, env_ids[:, None], ep[:, None], np.array([100, 101], dtype=np.uint64)[None, :])
    qpos[:, tgt] += -0.01 + 0.02 * u
    return qpos, np.zeros((n, model.nv))


class VecSawyerPushObstacle:
    """N device-resident SawyerPushObstacle-v0 environments."""
    ENV_ID = "SawyerPushObstacle-v0"
    OBS_DIM = 40
    INIT_QPOS = PUSH_INIT_QPOS
    STATIC_BODIES = ("table", "bin1")          # SawyerPushObstacleEnv.static_bodies
    MANIPULATION_BODIES = ("cube",)            # bodies whose geoms may touch the static ones (manipulation_geom_ids)
    make_task = staticmethod(make_push_task)
    reset_state = staticmethod(push_reset_state)

    def __init__(self, n_envs, seed=1234, device=0, env_id_offset=0, model=None, max_episode_steps=250, contacts=True, **task_kwargs):
        import torch

        self.torch = torch
        self.n = int(n_envs)
        self.seed = int(seed)
        self.model = model if model is not None else load_model(self.ENV_ID)
        self.dyn = DynModel(self.model)
        self.task = self.make_task(self.model, self.dyn, max_episode_steps=max_episode_steps, **task_kwargs)
        self.dev = torch.device("cuda", device)
        self.device_index = device
        L = lib()
        L.mopa_env_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
        L.mopa_env_destroy.argtypes = [C.c_void_p]
        L.mopa_env_destroy.restype = None
        L.mopa_env_enable_contacts.argtypes = [C.c_void_p, C.c_int32]
        L.mopa_env_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.mopa_env_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        self._L = L
        h = C.c_void_p()
        check(L.mopa_env_create(C.byref(self.dyn.desc), C.byref(self.task), int(device), C.byref(h)))
        self.h = h
        if not contacts:
            check(L.mopa_env_enable_contacts(self.h, 0))
        m, n = self.model, self.n
        f64, dev = torch.float64, self.dev
        self.qpos = torch.zeros(n, m.nq, dtype=f64, device=dev)
        self.qvel = torch.zeros(n, m.nv, dtype=f64, device=dev)
        self.prev_state = torch.zeros(n, 7, dtype=f64, device=dev)
        self.bias_prev = torch.zeros(n, BIAS_PAD, dtype=f64, device=dev)
        self.has_prev = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.ep_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self.ep_rew = torch.zeros(n, dtype=f64, device=dev)
        self.obs = torch.zeros(n, OBS_DIM, dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=f64, device=dev)
        self.done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.success = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.ncon = torch.zeros(n, dtype=torch.int32, device=dev)
        self.work = torch.zeros(n, dtype=torch.int32, device=dev)
        self.cforce = torch.zeros(n, dtype=f64, device=dev)   # get_contact_force() of every env after the latest step
        self.grasp = torch.zeros(n, dtype=torch.uint8, device=dev)      # lift: finger-touch flags of the latest simulated step
        self.unstable = torch.zeros(n, dtype=torch.uint8, device=dev)   # 1 = the latest step diverged and was discarded (episode ends)
        self.i_term = torch.zeros(n, 4, dtype=f64, device=dev)          # Pusher: integral term of the PID law
        self.buf = EnvBuffers(*[getattr(self, k).data_ptr() for k, _ in EnvBuffers._fields_])
        self.env_ids = np.arange(n, dtype=np.int64) + int(env_id_offset)
        self.episode_idx = np.zeros(n, dtype=np.int64)
        self.ref_joint_pos_indexes = [int(self.task.arm_qadr[k]) for k in range(7)]

    def close(self):
        if getattr(self, "h", None):
            self._L.mopa_env_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def forward(self, ids=None):
        """sim.forward(): refresh bias / observation of the given envs (torch int32 tensor) or all."""
        if ids is None:
            check(self._L.mopa_env_forward(self.h, C.byref(self.buf), None, self.n, C.c_void_p(self._stream())))
        elif len(ids):
            ids = ids.to(device=self.dev, dtype=self.torch.int32).contiguous()
            check(self._L.mopa_env_forward(self.h, C.byref(self.buf), C.c_void_p(ids.data_ptr()), int(len(ids)), C.c_void_p(self._stream())))
            self.torch.cuda.current_stream(self.dev).synchronize()  # ids must outlive the launch

    def reset(self, ids=None):
        """env.reset() for the listed env rows (numpy int array) or all; returns the observation tensor."""
        torch = self.torch
        ids = np.arange(self.n) if ids is None else np.asarray(ids, dtype=np.int64).reshape(-1)
        if len(ids) == 0:
            return self.obs
        qpos, qvel = self.reset_state(self.model, self.seed, self.env_ids[ids], self.episode_idx[ids])
        self.episode_idx[ids] += 1
        t_ids = torch.as_tensor(ids, device=self.dev)
        self.qpos[t_ids] = torch.as_tensor(qpos, device=self.dev)
        self.qvel[t_ids] = torch.as_tensor(qvel, device=self.dev)
        self.has_prev[t_ids] = 0
        self.ep_len[t_ids] = 0
        self.ep_rew[t_ids] = 0
        self.done[t_ids] = 0
        self.success[t_ids] = 0
        self.grasp[t_ids] = 0
        self.unstable[t_ids] = 0
        self.forward(t_ids)
        return self.obs

    def set_state(self, ids, qpos, qvel):
        """BaseEnv.set_state(qpos, qvel) + sim.forward() (env/base.py:419-427) for the listed env rows."""
        torch = self.torch
        t_ids = torch.as_tensor(np.asarray(ids, dtype=np.int64).reshape(-1), device=self.dev)
        self.qpos[t_ids] = torch.as_tensor(np.atleast_2d(qpos), dtype=torch.float64, device=self.dev)
        self.qvel[t_ids] = torch.as_tensor(np.atleast_2d(qvel), dtype=torch.float64, device=self.dev)
        self.forward(t_ids)

    def step(self, action, is_planner=None, mask=None):
        """env.step for all envs (or those with mask != 0).  action: float32 cuda tensor [n, >=7];
        is_planner / mask: uint8 cuda tensors [n] or None.  Results land in self.obs/reward/done/success."""
        a = action if action.dtype == self.torch.float32 else action.float()
        a = a.contiguous()
        check(self._L.mopa_env_step(self.h, C.byref(self.buf), C.c_void_p(a.data_ptr()), int(a.shape[1]),
                                    C.c_void_p(is_planner.data_ptr()) if is_planner is not None else None,
                                    C.c_void_p(mask.data_ptr()) if mask is not None else None, self.n, C.c_void_p(self._stream())))
        self._keep = (a, is_planner, mask)
        return self.obs, self.reward, self.done

    def reset_prev_state(self, mask=None):
        """env._reset_prev_state() (env/base.py:228-229)."""
        if mask is None:
            self.has_prev.zero_()
        else:
            self.has_prev[mask.bool()] = 0


class VecSawyerAssemblyObstacle(VecSawyerPushObstacle):
    """N device-resident SawyerAssemblyObstacle-v0 environments (BASELINE configs[3]).  The observation row keeps the
    40-float stride; its first 38 floats are joint_pos7, joint_vel7, gripper_qpos2, gripper_qvel2, eef_pos3, eef_quat4,
    hole3, pegHead3, pegEnd3, peg_quat4 (env/sawyer/sawyer_assembly_obstacle.py:53-59)."""
    ENV_ID = "SawyerAssemblyObstacle-v0"
    OBS_DIM = 38
    INIT_QPOS = ASSEMBLY_INIT_QPOS
    STATIC_BODIES = ("table",)                 # sawyer_assembly_obstacle.py:61-67
    MANIPULATION_BODIES = ("furniture", "0_part0", "1_part1", "4_part4", "2_part2")
    make_task = staticmethod(make_assembly_task)
    reset_state = staticmethod(assembly_reset_state)


class VecSawyerLiftObstacle(VecSawyerPushObstacle):
    """N device-resident SawyerLiftObstacle-v0 environments (BASELINE configs[2]).  Actions are 8-D (7 joint entries +
    gripper); the observation row keeps the 40-float stride, its first 35 floats are joint_pos7, joint_vel7, gripper_qpos2,
    gripper_qvel2, eef_pos3, eef_quat4, cube_pos3, cube_quat4, gripper_to_cube3 (env/sawyer/sawyer_lift_obstacle.py:150-161)."""
    ENV_ID = "SawyerLiftObstacle-v0"
    OBS_DIM = 35
    ACTION_DIM = 8
    INIT_QPOS = LIFT_INIT_QPOS
    STATIC_BODIES = ("table", "bin1")          # sawyer_lift_obstacle.py:163-165
    MANIPULATION_BODIES = ("cube",)
    make_task = staticmethod(make_lift_task)
    reset_state = staticmethod(lift_reset_state)


class VecPusherObstacle(VecSawyerPushObstacle):
    """N device-resident PusherObstacle-v0 environments (BASELINE configs[0]).  Actions are 4-D joint displacements; the
    observation row keeps the 40-float stride, its first 20 floats are cos / sin of the joint angles (4 + 4), box qpos 2, joint
    velocities 4, box velocity 2, fingertip xy 2, goal 2 (env/pusher/pusher_obstacle.py:185-205).  reset() runs the reference's
    rejection sampling (:40-67) on the device."""
    ENV_ID = "PusherObstacle-v0"
    OBS_DIM = 20
    ACTION_DIM = 4
    INIT_QPOS = np.zeros(7)
    STATIC_BODIES = ()
    MANIPULATION_BODIES = ()
    make_task = staticmethod(make_pusher_task)

    @staticmethod
    def planner_inputs(model):
        """ignored contacts / passive joints as rl/trainer.py:62-75 derives them from env/pusher/pusher_obstacle.py:
        manipulation geom `box` x the static obstacle geoms (:138-151), joints 0-3 active."""
        static = [model.geom_name2id("obstacle%d_geom" % i) for i in range(1, 8)]
        box = model.geom_name2id("box")
        ignored = [(min(box, g), max(box, g)) for g in static]
        ref = [model.get_joint_qpos_addr("joint%d" % i) for i in range(4)]
        return ignored, [i for i in range(model.nq) if i not in ref], ref

    def __init__(self, n_envs, seed=1234, device=0, env_id_offset=0, model=None, max_episode_steps=400, contacts=True, **task_kwargs):
        super().__init__(n_envs, seed=seed, device=device, env_id_offset=env_id_offset, model=model, max_episode_steps=max_episode_steps,
                         contacts=contacts, **task_kwargs)
        self.ref_joint_pos_indexes = [int(self.task.arm_qadr[k]) for k in range(4)]
        self.episode_dev = self.torch.zeros(self.n, dtype=self.torch.int64, device=self.dev)   # episode number of every env (reset draws)
        self._qpos0 = np.ascontiguousarray(self.model.qpos0, dtype=np.float64)
        self._L.mopa_env_reset_pusher.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]

    def reset(self, ids=None):
        torch = self.torch
        mask = None
        if ids is not None:
            ids = np.asarray(ids, dtype=np.int64).reshape(-1)
            if len(ids) == 0:
                return self.obs
            mask = torch.zeros(self.n, dtype=torch.uint8, device=self.dev)
            mask[torch.as_tensor(ids, device=self.dev)] = 1
        check(self._L.mopa_env_reset_pusher(self.h, C.byref(self.buf), C.c_void_p(mask.data_ptr()) if mask is not None else None,
                                            int(self.seed) & 0xFFFFFFFFFFFFFFFF, int(self.env_ids[0]), C.c_void_p(self.episode_dev.data_ptr()),
                                            self._qpos0.ctypes.data_as(C.c_void_p), self.n, C.c_void_p(self._stream())))
        self._keep_mask = mask
        self.episode_idx = self.episode_dev.cpu().numpy()
        return self.obs
