"""The reference's gym-style env API on top of the device-resident vectorised environments.

``rl/trainer.py:49`` builds its env with ``gym.make(config.env, **config.__dict__)`` (registration in
``env/__init__.py:7-32``, gym 0.15.4) and ``rl/mopa_rollouts.py`` / ``rl/sac_agent.py`` then use the
``BaseEnv`` / ``SawyerEnv`` surface listed in SURVEY.md section 8(b): ``reset / step(action, is_planner) / seed``,
``observation_space / action_space / joint_space``, ``sim.data.qpos / qvel / ncon / get_site_xpos``,
``sim.model.nq / nu / jnt_limited``, ``form_action``, ``compute_reward`` + ``_after_step`` (planner-failure step),
``set_state``, ``_reset_prev_state``, ``get_contact_force`` ...  gym itself is not installed here, so the few
pieces of it the reference touches (``spaces.Box / Dict / Discrete``, ``Env``, ``register``, ``make``) are part of
this module.

``make("SawyerPushObstacle-v0", **kwargs)`` returns an N = 1 view: every call goes to the same CUDA kernels the
vectorised runner uses (``mopa_env_step`` / ``mopa_env_forward`` through ``VecSawyer*``); nothing is computed on the
host except the kinematics behind ``get_site_xpos / get_site_xmat`` of sites the observation does not already carry.
The batched runner (``rollout.NativeMoPARolloutRunner``) is the fast path; this class is the drop-in one.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from .mjcf import quat_mul, quat_to_mat


# ------------------------------------------------------------------------------------ gym shim
class Space:
    shape = ()


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
            shape = low.shape
        else:
            low, high = np.full(shape, low, dtype=dtype), np.full(shape, high, dtype=dtype)
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Discrete(Space):
    def __init__(self, n):
        self.n, self.shape, self._rng = int(n), (), np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return int(self._rng.integers(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = OrderedDict(spaces)

    def seed(self, seed=None):
        for k, s in enumerate(self.spaces.values()):
            s.seed(None if seed is None else seed + k)

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def __getitem__(self, k):
        return self.spaces[k]


class spaces:  # ``from gym import spaces`` look-alike
    Box, Dict, Discrete = Box, Dict, Discrete


class Env:
    metadata = {}

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def seed(self, seed=None):
        return [seed]

    def close(self):
        pass


_REGISTRY = {}


def register(id, entry_point, kwargs=None):
    _REGISTRY[id] = (entry_point, dict(kwargs or {}))


def make(id, **kwargs):
    """gym.make: registered defaults updated by the caller's keyword arguments (rl/trainer.py:49)."""
    if id not in _REGISTRY:
        raise KeyError("No registered env with id: %s" % id)
    entry, defaults = _REGISTRY[id]
    kw = dict(defaults)
    kw.update(kwargs)
    return entry(**kw)


# ------------------------------------------------------------------------------------ host kinematics (sites)
def body_frames(model, qpos):
    """mj_kinematics for all bodies (pos, quat wxyz): used only behind get_site_xpos / get_site_xmat."""
    m, q = model, np.asarray(qpos, np.float64)
    pos, quat = np.zeros((m.nbody, 3)), np.tile(np.array([1.0, 0, 0, 0]), (m.nbody, 1))
    for b in range(1, m.nbody):
        p = int(m.body_parentid[b])
        bp = pos[p] + quat_to_mat(quat[p]) @ m.body_pos[b]
        bq = quat_mul(quat[p], m.body_quat[b])
        for j in range(int(m.body_jntadr[b]), int(m.body_jntadr[b]) + int(m.body_jntnum[b])):
            t, a = int(m.jnt_type[j]), int(m.jnt_qposadr[j])
            if t == 0:
                bp, bq = q[a:a + 3].copy(), q[a + 3:a + 7] / np.linalg.norm(q[a + 3:a + 7])
            elif t == 3:
                anchor = bp + quat_to_mat(bq) @ m.jnt_pos[j]
                ang = q[a] - m.qpos0[a]
                bq = quat_mul(bq, np.concatenate([[np.cos(0.5 * ang)], np.sin(0.5 * ang) * m.jnt_axis[j]]))
                bp = anchor - quat_to_mat(bq) @ m.jnt_pos[j]
            elif t == 2:
                bp = bp + quat_to_mat(bq) @ m.jnt_axis[j] * (q[a] - m.qpos0[a])
            else:
                raise NotImplementedError("ball joints")
        pos[b], quat[b] = bp, bq
    return pos, quat


class _SimData:
    """The slice of mujoco_py's sim.data the rollout code reads."""

    def __init__(self, env):
        self._env = env

    @property
    def qpos(self):
        return self._env._qpos_host()

    @property
    def qvel(self):
        return self._env._venv.qvel[0].cpu().numpy()

    @property
    def ncon(self):
        return int(self._env._venv.ncon[0].item())

    def _site(self, name):
        env = self._env
        sid = env.model.site_name2id(name)
        pos, quat = body_frames(env.model, self.qpos)
        b = int(env.model.site_bodyid[sid])
        R = quat_to_mat(quat[b])
        return pos[b] + R @ env.model.site_pos[sid], R @ quat_to_mat(env.model.site_quat[sid])

    def get_site_xpos(self, name):
        """Site position at the CURRENT qpos (what mjData holds after sim.forward()); mujoco-py returns the frames of the
        last mj_step's start state when no forward() was called since."""
        return self._site(name)[0]

    def get_site_xmat(self, name):
        return self._site(name)[1]

    def get_body_xpos(self, name):
        return body_frames(self._env.model, self.qpos)[0][self._env.model.body_name2id(name)]


class _Sim:
    def __init__(self, env):
        self.model, self.data = env.model, _SimData(env)
        self._env = env

    def forward(self):
        self._env._venv.forward()

    def get_state(self):
        return self.data.qpos.copy(), self.data.qvel.copy()


# ------------------------------------------------------------------------------------ the env view
_OBS_KEYS = {
    "push": (("joint_pos", 7), ("joint_vel", 7), ("gripper_qpos", 2), ("gripper_qvel", 2), ("eef_pos", 3), ("eef_quat", 4),
             ("target_pos", 3), ("cube_pos", 3), ("cube_quat", 4), ("gripper_to_cube", 3), ("cube_to_target", 2)),
    "lift": (("joint_pos", 7), ("joint_vel", 7), ("gripper_qpos", 2), ("gripper_qvel", 2), ("eef_pos", 3), ("eef_quat", 4),
             ("cube_pos", 3), ("cube_quat", 4), ("gripper_to_cube", 3)),
    "assembly": (("joint_pos", 7), ("joint_vel", 7), ("gripper_qpos", 2), ("gripper_qvel", 2), ("eef_pos", 3), ("eef_quat", 4),
                 ("hole", 3), ("pegHead", 3), ("pegEnd", 3), ("peg_quat", 4)),   # sawyer_assembly_obstacle.py:52-58
    "pusher": (("default", 16), ("fingertip", 2), ("goal", 2)),                   # pusher_obstacle.py:185-205
}


class SawyerEnvView(Env):
    """One environment with the reference's ``SawyerEnv`` surface (env/base.py, env/sawyer/sawyer.py)."""
    TASK = "push"
    VEC_CLASS = "VecSawyerPushObstacle"
    ROBOT_JOINTS = tuple("right_j%d" % i for i in range(7))
    GRIPPER_JOINTS = ("rc_close", "lc_close")
    WORLD = ([-1.2, -1.2, 0.0], [1.2, 1.2, 2.0])   # env/sawyer/sawyer.py:52-53

    def __init__(self, seed=1234, device=0, max_episode_steps=250, venv=None, **kwargs):
        from . import envs

        self._kwargs = dict(kwargs)
        task_kw = {k: kwargs[k] for k in ("frame_dt", "ac_scale", "distance_threshold", "success_reward") if k in kwargs and
                   not (k == "distance_threshold" and self.TASK != "push")}
        self._venv = venv if venv is not None else getattr(envs, self.VEC_CLASS)(1, seed=int(seed), device=device,
                                                                                   max_episode_steps=max_episode_steps, **task_kw)
        self.model = self._venv.model
        m = self.model
        self.sim = _Sim(self)
        self.data = self.sim.data
        # the reference hands env.xml_path to the planner (rl/trainer.py:56); model.load_model resolves this name to the
        # MJCF file when the reference's asset tree is reachable, else to the cached compiled scene next to it
        from .model import _ALIASES, ASSET_DIR
        import os

        self.xml_path = os.path.join(ASSET_DIR, _ALIASES[getattr(envs, self.VEC_CLASS).ENV_ID] + ".xml")
        self.max_episode_steps = int(max_episode_steps)
        self._ac_scale = float(self._venv.task.ac_scale)
        self.robot_joints = list(self.ROBOT_JOINTS)
        self.ref_joint_pos_indexes = [m.get_joint_qpos_addr(j) for j in self.robot_joints]
        self.ref_joint_vel_indexes = [m.get_joint_qvel_addr(j) for j in self.robot_joints]
        self.ref_gripper_joint_pos_indexes = [m.get_joint_qpos_addr(j) for j in self.GRIPPER_JOINTS]
        self.dof = int(getattr(self._venv, "ACTION_DIM", 7))
        self.robot_dof = len(self.robot_joints)
        self.min_world_size, self.max_world_size = list(self.WORLD[0]), list(self.WORLD[1])
        # env/base.py:67-99: per-qpos joint index table, limits (unlimited -> +-3.14), joint_space
        self.jnt_indices = []
        for i, t in enumerate(m.jnt_type):
            self.jnt_indices += [i] * (7 if t == 0 else (4 if t == 1 else 1))
        lim = np.asarray(m.jnt_limited).astype(bool)
        self._is_jnt_limited = lim
        self._jnt_minimum = np.where(lim, m.jnt_range[:, 0], -3.14)
        self._jnt_maximum = np.where(lim, m.jnt_range[:, 1], 3.14)
        self.joint_space = Dict([("default", Box(low=self._jnt_minimum, high=self._jnt_maximum, dtype=np.float32))])
        self.action_space = Dict([("default", Box(-1.0, 1.0, shape=(self.dof,), dtype=np.float32))])
        self._keys = _OBS_KEYS[self.TASK]
        self.observation_space = Dict([(k, Box(-1.0, 1.0, shape=(n,), dtype=np.float32)) for k, n in self._keys])
        body_of = lambda g: m.names["body"][m.geom_bodyid[g]]
        cls = getattr(envs, self.VEC_CLASS)
        self.static_geom_ids = [g for g in range(m.ngeom) if body_of(g) in cls.STATIC_BODIES]
        self.manipulation_geom_ids = [g for g in range(m.ngeom) if body_of(g) in cls.MANIPULATION_BODIES]
        if hasattr(cls, "planner_inputs"):   # envs that name geoms instead of bodies (PusherObstacleEnv.static_geoms / manipulation_geom)
            ign, _, _ = cls.planner_inputs(m)
            self.manipulation_geom_ids = sorted({a for a, b in ign} & {m.geom_name2id("box")}) or sorted({a for a, _ in ign})
            self.static_geom_ids = sorted({g for pair in ign for g in pair} - set(self.manipulation_geom_ids))
        self._pending = None       # (done, info) of a compute_reward() call waiting for its _after_step()
        self._last = (0.0, False)

    # ---- helpers
    def _qpos_host(self):
        return self._venv.qpos[0].cpu().numpy()

    def _ob(self):
        row = self._venv.obs[0].cpu().numpy()
        out, o = OrderedDict(), 0
        for k, n in self._keys:
            out[k] = row[o:o + n].astype(np.float64)
            o += n
        return out

    def _info(self, done):
        info = {}
        if done:
            info = dict(episode_success=int(self._venv.success[0].item()), episode_reward=float(self._venv.ep_rew[0].item()),
                        episode_length=int(self._venv.ep_len[0].item()),
                        episode_unstable=int(self._venv.unstable[0].item()) if hasattr(self._venv, "unstable") else 0)
        return info

    # ---- gym API
    def seed(self, seed=None):
        if seed is not None:
            self._venv.seed = int(seed)
        return [self._venv.seed]

    def reset(self):
        self._venv.reset()
        self._pending = None
        return self._ob()

    def step(self, action, is_planner=False):
        """BaseEnv.step (env/base.py:232-247): dict / list / array actions, 4-tuple result."""
        if isinstance(action, list):
            action = {key: val for ac_i in action for key, val in ac_i.items()}
        if isinstance(action, dict):
            action = np.concatenate([np.atleast_1d(action[k]) for k in self.action_space.spaces.keys() if k != "ac_type"])
        action = np.asarray(action, np.float64).reshape(-1)
        if len(action) != self.dof:
            raise AssertionError("environment got invalid action dimension")
        return self._native_step(action, 1 if is_planner else 0)

    def _native_step(self, action, mode):
        torch = self._venv.torch
        a = np.zeros((1, 8), np.float32)
        a[0, :len(action)] = action
        self._venv.step(torch.as_tensor(a, device=self._venv.dev), torch.full((1,), mode, dtype=torch.uint8, device=self._venv.dev))
        reward, done = float(self._venv.reward[0].item()), bool(self._venv.done[0].item())
        self._last = (reward, done)
        return self._ob(), reward, done, self._info(done)

    # ---- what rl/mopa_rollouts.py touches besides step()
    @property
    def _terminal(self):
        return bool(self._venv.done[0].item())

    @property
    def _success(self):
        return bool(self._venv.success[0].item())

    @property
    def _episode_length(self):
        return int(self._venv.ep_len[0].item())

    @property
    def _episode_reward(self):
        return float(self._venv.ep_rew[0].item())

    def _reset_prev_state(self):
        self._venv.reset_prev_state()

    def set_state(self, qpos, qvel):
        self._venv.set_state([0], np.asarray(qpos, np.float64)[None], np.asarray(qvel, np.float64)[None])

    def get_contact_force(self):
        return float(self._venv.cforce[0].item())

    def form_action(self, next_qpos, curr_qpos=None):
        """SawyerEnv.form_action (env/sawyer/sawyer.py:283-299)."""
        if curr_qpos is None:
            curr_qpos = self._qpos_host()
        next_qpos, curr_qpos = np.asarray(next_qpos), np.asarray(curr_qpos)
        joint_ac = next_qpos[self.ref_joint_pos_indexes] - curr_qpos[self.ref_joint_pos_indexes]
        if self.dof == 8:
            g = next_qpos[self.ref_gripper_joint_pos_indexes] - curr_qpos[self.ref_gripper_joint_pos_indexes]
            return OrderedDict([("default", np.concatenate([joint_ac, [g[0]]]))])
        return OrderedDict([("default", joint_ac)])

    def compute_reward(self, action):
        """Planner-failure step, first half (rl/mopa_rollouts.py:304-327 calls compute_reward(zeros) and then
        _after_step): both halves run in one device pass (mode 2 of mopa_env_step: reward + _after_step, no simulation);
        the second half is handed out by _after_step()."""
        _, reward, done, info = self._native_step(np.zeros(self.dof), 2)
        self._pending = (done, info)
        return reward, {}

    def _after_step(self, reward, terminal, info):
        if self._pending is None:
            raise RuntimeError("_after_step() is only available after compute_reward() (env.step runs it on the device)")
        done, step_log = self._pending
        self._pending = None
        merged = dict(info or {})
        merged.update(step_log)
        return done, merged, 0

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the scope of this package (DESIGN.md section 9)")

    # indicator / colour helpers of the reference are render-only
    def visualize_goal_indicator(self, *a, **k):
        pass

    visualize_dummy_indicator = reset_visualized_indicator = color_agent = reset_color_agent = visualize_goal_indicator

    def close(self):
        self._venv.close()


class SawyerPushObstacleEnv(SawyerEnvView):
    TASK, VEC_CLASS = "push", "VecSawyerPushObstacle"


class SawyerLiftObstacleEnv(SawyerEnvView):
    TASK, VEC_CLASS = "lift", "VecSawyerLiftObstacle"


class SawyerAssemblyObstacleEnv(SawyerEnvView):
    TASK, VEC_CLASS = "assembly", "VecSawyerAssemblyObstacle"


class PusherObstacleEnv(SawyerEnvView):
    """PusherObstacle-v0 (env/pusher/pusher_obstacle.py; BASELINE configs[0]): 4-D joint-displacement actions, observation
    keys default (cos / sin of the joint angles, box qpos, joint velocities, box velocity) / fingertip / goal."""
    TASK, VEC_CLASS = "pusher", "VecPusherObstacle"
    ROBOT_JOINTS = ("joint0", "joint1", "joint2", "joint3")
    GRIPPER_JOINTS = ()
    WORLD = ([-0.41, -0.41], [0.41, 0.41])        # pusher_obstacle.py:34-35

    def form_action(self, next_qpos, curr_qpos=None):
        """BaseEnv.form_action (env/base.py:402-410)."""
        if curr_qpos is None:
            curr_qpos = self._qpos_host()
        next_qpos, curr_qpos = np.asarray(next_qpos), np.asarray(curr_qpos)
        return OrderedDict([("default", next_qpos[self.ref_joint_pos_indexes] - curr_qpos[self.ref_joint_pos_indexes])])


# env/__init__.py:7-32
register("PusherObstacle-v0", PusherObstacleEnv)
register("SawyerPushObstacle-v0", SawyerPushObstacleEnv)
register("SawyerLiftObstacle-v0", SawyerLiftObstacleEnv)
register("SawyerAssemblyObstacle-v0", SawyerAssemblyObstacleEnv)
