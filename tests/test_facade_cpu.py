"""Host logic of the drop-in facades (no GPU): sentinel decoding, path re-basing, wrap handling —
written to read like checks of motion_planners/sampling_based_planner.py and rl/planner_agent.py."""
import types

import numpy as np
import pytest

from mopa_rl_b200.motion_planners import sampling_based_planner as sbp
from mopa_rl_b200.motion_planners.sampling_based_planner import SamplingBasedPlanner, joint_convert


class FakeNative:
    """Stands in for PyKinematicPlanner so that the Python-side logic can run without a GPU."""

    def __init__(self, *a, **k):
        self.calls = []
        self.next = None

    def plan(self, start, goal, timelimit):
        self.calls.append((np.array(start), np.array(goal), timelimit))
        return self.next

    def isValidState(self, s):
        return True

    def getPlannerStatus(self):
        return b"Exact solution"


@pytest.fixture
def planner(monkeypatch):
    monkeypatch.setattr(sbp, "PyKinematicPlanner", FakeNative)
    cfg = types.SimpleNamespace(planner_type="rrt_connect", range=0.1, planner_objective="path_length", threshold=0.0, seed=1)
    return lambda non_limited: SamplingBasedPlanner(cfg, "scene.xml", 7, non_limited, passive_joint_idx=[7, 8], ignored_contacts=[(1, 2)],
                                                    contact_threshold=-0.002)


def test_joint_convert_matches_reference_definition():
    # util/env.py:15-25 spelled out for a few hand-computed points (period 3.14, not pi)
    assert joint_convert(1.0) == 1.0 and np.isclose(joint_convert(4.0), 4.0 % 3.14 - 3.14)
    assert np.isclose(joint_convert(7.0), 7.0 % 3.14) and np.isclose(joint_convert(-4.0), -4.0 % -3.14 + 3.14)
    assert joint_convert(-1.0) == -1.0 and joint_convert(0.0) == 0.0
    for x in np.linspace(-20, 20, 401):
        y = joint_convert(x)
        assert -3.14 <= y <= 3.14 and np.isclose(np.cos(x * np.pi / 3.14), np.cos(y * np.pi / 3.14), atol=1e-9)


def test_sentinel_rows(planner):
    p = planner(None)
    start, goal = np.zeros(9), np.ones(9)
    p.planner.next = [[-5.0] * 9]
    traj, states, valid, exact = p.plan(start, goal, 1.0)
    assert traj.shape == (1, 9) and (traj == -5).all() and not valid and exact          # invalid goal
    p.planner.next = [[-4.0] * 9]
    traj, states, valid, exact = p.plan(start, goal, 1.0)
    assert (traj == -4).all() and valid and not exact                                   # no exact solution
    assert p.get_planner_status() == "Exact solution"


def test_path_is_rebased_on_the_unwrapped_start_and_inputs_are_not_mutated(planner):
    p = planner(None)
    start = np.array([0.1, 0.2, 0.3, 0, 0, 0, 0, 0.5, 0.6])
    goal = start + 0.2
    s0, g0 = start.copy(), goal.copy()
    states = np.array([start.astype(np.float32), start + 0.1, goal]).astype(np.float64)   # planner rows: fp32 start
    p.planner.next = states.tolist()
    traj, st, valid, exact = p.plan(start, goal, 2.0)
    assert valid and exact and traj.shape == (3, 9)
    assert np.array_equal(traj[0], start)                                                # exact start, not its fp32 image
    assert np.allclose(traj[1:] - traj[:-1], states[1:] - states[:-1])
    assert np.array_equal(start, s0) and np.array_equal(goal, g0)
    assert p.planner.calls[-1][2] == 2.0


def test_unlimited_joint_wrap_repair(planner):
    p = planner([0])
    start = np.array([3.0 + 3.14 * 2, 0, 0, 0, 0, 0, 0, 0, 0])      # wrapped to 3.0 before planning
    goal = np.array([-3.0, 0, 0, 0, 0, 0, 0, 0, 0])
    p.planner.next = [[3.0] + [0] * 8, [3.13] + [0] * 8, [-3.13] + [0] * 8, [-3.0] + [0] * 8]   # crosses the +-3.14 seam
    traj, _, _, _ = p.plan(start, goal, 1.0)
    assert np.isclose(p.planner.calls[-1][0][0], joint_convert(start[0]))
    d = np.diff(traj[:, 0])
    assert np.allclose(d, [0.13, 0.02, 0.13])                        # short way round, accumulated on the unwrapped start
    assert np.isclose(traj[0, 0], start[0])


def test_planner_agent_drops_first_waypoint(monkeypatch):
    from mopa_rl_b200 import planner_agent as pa

    class FakeSBP:
        def __init__(self, *a, **k):
            self.kw = k

        def plan(self, s, g, timelimit):
            self.timelimit = timelimit
            return (np.arange(12.0).reshape(4, 3), None, True, True) if g[0] >= 0 else (np.full((1, 3), -4.0), None, True, False)

        def isValidState(self, s):
            return s[0] > 0

        def get_planner_status(self):
            return "x"

    monkeypatch.setattr(pa, "SamplingBasedPlanner", FakeSBP)
    cfg = types.SimpleNamespace(_xml_path="scene.xml", contact_threshold=-0.002, timelimit=1.5)
    space = types.SimpleNamespace(spaces={"default": types.SimpleNamespace(shape=(7,))})
    agent = pa.PlannerAgent(cfg, space, non_limited_idx=None, passive_joint_idx=[1], ignored_contacts=[(3, 4)], range_=0.05)
    traj, success, valid, exact = agent.plan(np.zeros(3), np.ones(3))
    assert success and traj.shape == (3, 3) and traj[0, 0] == 3.0 and agent.planner.timelimit == 1.5
    traj, success, valid, exact = agent.plan(np.zeros(3), -np.ones(3), timelimit=0.1)
    assert not success and valid and not exact and traj.shape == (1, 3)
    assert agent.isValidState(np.ones(3)) and agent.planner.kw["range_"] == 0.05 and agent.planner.kw["contact_threshold"] == -0.002


def test_pykinematicplanner_argument_checks():
    from mopa_rl_b200.motion_planners.planner import PyKinematicPlanner

    args = dict(num_actions=7, opt=b"path_length", threshold=0.0, _range=0.1, passive_joint_idx=[], glue_bodies=[], ignored_contacts=[],
                contact_threshold=0.0, goal_bias=0.05, is_simplified=False, simplified_duration=0.1, seed=1)
    with pytest.raises(ValueError):
        PyKinematicPlanner(b"scene", b"rrt_connect", **args)               # "XML model file is required"
    with pytest.raises(NotImplementedError):
        PyKinematicPlanner(b"scene.xml", b"rrt", **args)                   # RRT* is not the reference default and is not built
