"""Host logic of the gym-API env view (mopa_rl_b200/gym_env.py) with a fake vectorised env: spaces, action
marshalling of BaseEnv.step (env/base.py:232-247), form_action, the compute_reward / _after_step protocol of the
planner-failure step, joint tables (env/base.py:67-99), and the host kinematics behind get_site_xpos against the
oracle's forward kinematics."""
from collections import OrderedDict

import numpy as np
import pytest


class _Task:
    ac_scale = 0.05


class FakeVenv:
    """Records what the view sends to the device layer; CPU torch tensors stand in for the device arrays."""

    def __init__(self, model, action_dim=7):
        import torch

        self.torch, self.dev, self.model, self.task, self.seed = torch, torch.device("cpu"), model, _Task(), 0
        self.ACTION_DIM = action_dim
        self.qpos = torch.as_tensor(model.qpos0[None].copy())
        self.qvel = torch.zeros(1, model.nv, dtype=torch.float64)
        self.obs = torch.arange(40, dtype=torch.float32)[None].clone()
        self.reward, self.ep_rew = torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
        self.done, self.success = torch.zeros(1, dtype=torch.uint8), torch.zeros(1, dtype=torch.uint8)
        self.ncon, self.ep_len = torch.zeros(1, dtype=torch.int32), torch.zeros(1, dtype=torch.int32)
        self.cforce = torch.full((1,), 2.5, dtype=torch.float64)
        self.calls = []

    def reset(self):
        self.calls.append(("reset",))

    def forward(self):
        self.calls.append(("forward",))

    def step(self, action, is_planner=None, mask=None):
        self.calls.append(("step", action.numpy().copy(), int(is_planner[0])))
        self.reward[0] = 0.25
        self.ep_len += 1
        self.ep_rew += 0.25
        self.done[0] = 1 if int(self.ep_len[0]) >= 3 else 0

    def set_state(self, ids, qpos, qvel):
        self.calls.append(("set_state", np.array(qpos), np.array(qvel)))
        self.qpos[0] = self.torch.as_tensor(qpos[0])

    def reset_prev_state(self, mask=None):
        self.calls.append(("reset_prev_state",))

    def close(self):
        pass


def test_gym_shim_and_registration():
    from mopa_rl_b200 import gym_env as G

    b = G.Box(-1.0, 1.0, shape=(7,), dtype=np.float32)
    assert b.shape == (7,) and b.contains(b.sample()) and not b.contains(np.full(7, 2.0, np.float32))
    d = G.Dict([("default", b), ("ac_type", G.Discrete(2))])
    s = d.sample()
    assert list(s.keys()) == ["default", "ac_type"] and s["ac_type"] in (0, 1)
    assert set(G._REGISTRY) >= {"SawyerPushObstacle-v0", "SawyerLiftObstacle-v0", "SawyerAssemblyObstacle-v0", "PusherObstacle-v0"}
    with pytest.raises(KeyError):
        G.make("Nope-v0")


@pytest.mark.parametrize("cls_name,scene,dof,obs_dim", [("SawyerPushObstacleEnv", "SawyerPushObstacle-v0", 7, 40),
                                                         ("SawyerLiftObstacleEnv", "SawyerLiftObstacle-v0", 8, 35),
                                                         ("SawyerAssemblyObstacleEnv", "SawyerAssemblyObstacle-v0", 7, 38),
                                                         ("PusherObstacleEnv", "PusherObstacle-v0", 4, 20)])
def test_env_view_spaces_and_tables(cls_name, scene, dof, obs_dim):
    from mopa_rl_b200 import gym_env as G
    from mopa_rl_b200.model import load_model

    m = load_model(scene)
    env = getattr(G, cls_name)(venv=FakeVenv(m, dof), max_episode_steps=3)
    assert env.action_space["default"].shape == (dof,)
    assert sum(s.shape[0] for s in env.observation_space.spaces.values()) == obs_dim
    ob = env.reset()
    assert isinstance(ob, OrderedDict) and np.array_equal(np.concatenate(list(ob.values())), np.arange(obs_dim))
    # observation keys in the reference's order (env/sawyer/sawyer.py:317-338 + the task's _get_obs)
    pusher = cls_name == "PusherObstacleEnv"
    if pusher:   # env/pusher/pusher_obstacle.py:185-205
        assert list(ob.keys()) == ["default", "fingertip", "goal"] and [len(v) for v in ob.values()] == [16, 2, 2]
    else:
        tail = {"SawyerPushObstacleEnv": ["target_pos", "cube_pos", "cube_quat", "gripper_to_cube", "cube_to_target"],
                "SawyerLiftObstacleEnv": ["cube_pos", "cube_quat", "gripper_to_cube"],
                "SawyerAssemblyObstacleEnv": ["hole", "pegHead", "pegEnd", "peg_quat"]}[cls_name]
        assert list(ob.keys()) == ["joint_pos", "joint_vel", "gripper_qpos", "gripper_qvel", "eef_pos", "eef_quat"] + tail
    # env/base.py:67-99: one jnt_indices entry per qpos element, free joints count 7 times; unlimited joints +-3.14
    assert len(env.jnt_indices) == m.nq and env.sim.model.nq == m.nq
    lim = np.asarray(m.jnt_limited).astype(bool)
    assert np.all(env._jnt_minimum[~lim] == -3.14) and np.all(env._jnt_maximum[~lim] == 3.14)
    assert np.array_equal(env._jnt_minimum[lim], m.jnt_range[lim, 0])
    assert env.joint_space["default"].shape == (m.njnt,)
    if pusher:
        assert env.ref_joint_pos_indexes == list(range(4)) and not lim[0] and env.min_world_size == [-0.41, -0.41]
        assert [m.names["geom"][g] if "geom" in m.names else g for g in env.manipulation_geom_ids] and len(env.static_geom_ids) == 7
        assert env.manipulation_geom_ids == [m.geom_name2id("box")]
        assert np.allclose(env.form_action(m.qpos0 + 0.03)["default"], 0.03) and len(env.form_action(m.qpos0)["default"]) == 4
        return
    assert env.ref_joint_pos_indexes == list(range(7)) and env._ac_scale == 0.05
    cube_like = set(m.names["body"][m.geom_bodyid[g]] for g in env.manipulation_geom_ids)
    assert cube_like and all(m.names["body"][m.geom_bodyid[g]] in ("table", "bin1") for g in env.static_geom_ids)


def test_step_marshalling_and_failure_protocol(push_model):
    from mopa_rl_b200 import gym_env as G

    fv = FakeVenv(push_model)
    env = G.SawyerPushObstacleEnv(venv=fv, max_episode_steps=3)
    a = np.linspace(-1, 1, 7)
    ob, r, d, info = env.step(OrderedDict([("default", a), ("ac_type", np.array([1]))]))       # ac_type is not part of the env action
    assert fv.calls[-1][0] == "step" and np.allclose(fv.calls[-1][1][0, :7], a) and fv.calls[-1][2] == 0
    assert r == 0.25 and d is False and info == {}
    env.step([{"default": a * 0.01}], is_planner=True)                                         # list-of-dicts form
    assert fv.calls[-1][2] == 1 and np.allclose(fv.calls[-1][1][0, :7], a * 0.01)
    with pytest.raises(AssertionError):
        env.step(np.zeros(8))
    with pytest.raises(RuntimeError):
        env._after_step(0.0, False, {})
    reward, info = env.compute_reward(np.zeros(7))                                             # rl/mopa_rollouts.py:304-327
    assert fv.calls[-1][2] == 2 and np.all(fv.calls[-1][1] == 0) and reward == 0.25
    done, info, penalty = env._after_step(reward, False, info)
    assert done is True and penalty == 0 and info["episode_length"] == 3 and abs(info["episode_reward"] - 0.75) < 1e-12
    assert env._terminal and env._episode_length == 3 and env.get_contact_force() == 2.5
    env._reset_prev_state()
    assert fv.calls[-1] == ("reset_prev_state",)
    q = push_model.qpos0.copy()
    q[:7] += 0.1
    env.set_state(q, np.zeros(push_model.nv))
    assert np.allclose(env.sim.data.qpos, q)
    assert np.allclose(env.form_action(q + 0.02)["default"], 0.02)
    with pytest.raises(NotImplementedError):
        env.render()


def test_form_action_with_gripper_entry():
    from mopa_rl_b200 import gym_env as G
    from mopa_rl_b200.model import load_model

    m = load_model("SawyerLiftObstacle-v0")
    env = G.SawyerLiftObstacleEnv(venv=FakeVenv(m, 8))
    nxt = m.qpos0.copy()
    nxt[:7] += 0.03
    nxt[env.ref_gripper_joint_pos_indexes] += [0.004, -0.002]
    ac = env.form_action(nxt, m.qpos0)["default"]                                             # env/sawyer/sawyer.py:290-296
    assert ac.shape == (8,) and np.allclose(ac[:7], 0.03) and abs(ac[7] - 0.004) < 1e-15


def test_site_kinematics_match_oracle_fk(push_model, oracle_built):
    from helpers import planner_setup, random_qpos
    from mopa_rl_b200 import gym_env as G

    ignored, passive, ref = planner_setup(push_model)
    scene = oracle_built.OracleScene(push_model, ignored, -0.002, "f64")
    fv = FakeVenv(push_model)
    env = G.SawyerPushObstacleEnv(venv=fv)
    for q in random_qpos(push_model, 5, 3, ref):
        q = q.copy()
        q[27:34] = [0.7, 0.1, 1.0, 0.6, 0.0, 0.8, 0.0]                                         # cube free joint
        fv.qpos[0] = fv.torch.as_tensor(q)
        fk = scene.fk(q)
        for name in ("grip_site", "right_eef", "cube"):
            sid = push_model.site_name2id(name)
            assert np.abs(env.sim.data.get_site_xpos(name) - fk["site_xpos"][sid]).max() < 1e-12
            assert np.abs(env.sim.data.get_site_xmat(name).ravel() - fk["site_xmat"][sid]).max() < 1e-12


@pytest.mark.parametrize("scene", ["SawyerLiftObstacle-v0", "SawyerAssemblyObstacle-v0", "PusherObstacle-v0"])
def test_host_kinematics_match_oracle_fk_on_every_scene(scene, oracle_built):
    """gym_env.body_frames (hinge / slide / free joints, bodies with several joints: the Pusher's box and target) against
    the oracle's mj_kinematics restatement: body and site frames of random states."""
    from mopa_rl_b200 import gym_env as G
    from mopa_rl_b200.mjcf import quat_to_mat
    from mopa_rl_b200.model import load_model

    m = load_model(scene)
    scn = oracle_built.OracleScene(m, [], 0.0, "f64")
    rng = np.random.default_rng(8)
    for _ in range(4):
        q = m.qpos0 + rng.uniform(-0.3, 0.3, m.nq)
        for j in range(m.njnt):                      # free joints: a proper pose
            if m.jnt_type[j] == 0:
                a = int(m.jnt_qposadr[j])
                quat = rng.normal(size=4)
                q[a + 3:a + 7] = quat / np.linalg.norm(quat)
        pos, quat = G.body_frames(m, q)
        fk = scn.fk(q)
        assert np.abs(pos - fk["body_xpos"]).max() < 1e-12
        R = np.stack([quat_to_mat(x).ravel() for x in quat])
        assert np.abs(R - fk["body_xmat"]).max() < 1e-12
