"""``PlannerAgent`` with the reference's interface (``rl/planner_agent.py:8-58``)."""
from __future__ import annotations

import numpy as np

from .motion_planners import SamplingBasedPlanner


def _action_size(ac_space):
    if hasattr(ac_space, "spaces"):
        return int(sum(_action_size(v) for v in ac_space.spaces.values()))
    if hasattr(ac_space, "shape") and ac_space.shape is not None and len(ac_space.shape):
        return int(np.prod(ac_space.shape))
    return int(getattr(ac_space, "n", 1))


class PlannerAgent:
    def __init__(self, config, ac_space, non_limited_idx=None, passive_joint_idx=[], ignored_contacts=[], planner_type=None,
                 goal_bias=0.05, is_simplified=False, simplified_duration=0.1, range_=None):
        self._config = config
        self.planner = SamplingBasedPlanner(
            config, config._xml_path, _action_size(ac_space), non_limited_idx, planner_type=planner_type,
            passive_joint_idx=passive_joint_idx, ignored_contacts=ignored_contacts,
            contact_threshold=config.contact_threshold, goal_bias=goal_bias, is_simplified=is_simplified,
            simplified_duration=simplified_duration, range_=range_)
        self._is_simplified = is_simplified
        self._simplified_duration = simplified_duration

    def plan(self, start, goal, timelimit=None, attempts=15):
        if timelimit is None:
            timelimit = self._config.timelimit
        traj, states, valid, exact = self.planner.plan(start, goal, timelimit)
        success = valid and exact
        return (traj[1:] if success else traj), success, valid, exact

    def get_planner_status(self):
        return self.planner.get_planner_status()

    def isValidState(self, state):
        return self.planner.isValidState(state)
