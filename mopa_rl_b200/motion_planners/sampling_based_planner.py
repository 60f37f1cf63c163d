"""``SamplingBasedPlanner`` — the Python API boundary of the reference
(``motion_planners/sampling_based_planner.py:11-107``) over the B200 planner.

Same constructor signature, same ``plan`` / ``isValidState`` / ``get_planner_status`` /
``convert_nonlimited`` contract:

* unlimited joints are wrapped into (-3.14, 3.14) before planning (the reference's "pi" is
  3.14, ``util/env.py:15-25``);
* a failed plan comes back as the 1 x nq sentinel matrix (all -5: invalid goal ->
  ``valid_state=False``; all -4: no exact solution -> ``exact=False``), detected exactly like
  the reference does (one unique value);
* a successful path is re-based on the un-wrapped ``start`` by accumulating waypoint deltas,
  repairing +-3.14 wrap-arounds on unlimited joints.
"""
from __future__ import annotations

import numpy as np

from .planner import PyKinematicPlanner


def joint_convert(angle):
    """Wrap an angle the way ``util/env.py:15-25`` does (period 3.14, sign preserving)."""
    period = 3.14 if angle > 0 else -3.14
    wrapped = angle % period
    if (angle // period) % 2 != 0:
        wrapped -= period
    return wrapped


class SamplingBasedPlanner:
    def __init__(self, config, xml_path, num_actions, non_limited_idx, planner_type=None, passive_joint_idx=[],
                 glue_bodies=[], ignored_contacts=[], contact_threshold=0.0, goal_bias=0.05, is_simplified=False,
                 simplified_duration=0.1, range_=None):
        self.config = config
        planner_type = config.planner_type if planner_type is None else planner_type
        range_ = config.range if range_ is None else range_
        self.planner = PyKinematicPlanner(
            xml_path.encode("utf-8"), planner_type.encode("utf-8"), num_actions, config.planner_objective.encode("utf-8"),
            config.threshold, range_, passive_joint_idx, glue_bodies, ignored_contacts, contact_threshold, goal_bias,
            is_simplified, simplified_duration, config.seed, device=getattr(config, "planner_device", 0))
        self.non_limited_idx = non_limited_idx

    def convert_nonlimited(self, state):
        if self.non_limited_idx is not None:
            for idx in self.non_limited_idx:
                state[idx] = joint_convert(state[idx])
        return state

    def isValidState(self, state):
        return self.planner.isValidState(state)

    def plan(self, start, goal, timelimit=1.0):
        wrapped_start = self.convert_nonlimited(np.array(start, dtype=np.float64))
        wrapped_goal = self.convert_nonlimited(np.array(goal, dtype=np.float64))
        states = np.array(self.planner.plan(wrapped_start, wrapped_goal, timelimit))

        if np.unique(states).size == 1:  # sentinel row
            code = states[0][0]
            return states, states, code != -5, code != -4

        deltas = np.diff(states, axis=0)
        wrap = np.zeros_like(deltas)
        if self.non_limited_idx is not None:
            for idx in self.non_limited_idx:
                prev, cur = states[:-1, idx], states[1:, idx]
                jumped = np.abs(cur - prev) > 3.14
                # crossing +3.14 -> -3.14: the true motion is the short way round
                fwd = jumped & (prev > 0) & (cur <= 0)
                bwd = jumped & (prev < 0) & (cur > 0)
                wrap[fwd, idx] = (3.14 - prev[fwd] + cur[fwd] + 3.14) - deltas[fwd, idx]
                wrap[bwd, idx] = -(3.14 - cur[bwd] + prev[bwd] + 3.14) - deltas[bwd, idx]
        traj = np.vstack([np.asarray(start, dtype=np.float64)[None], start + np.cumsum(deltas + wrap, axis=0)])
        return traj, states, True, True

    def remove_collision(self, geom_id, contype, conaffinity):
        # the reference forwards to a binding that does not exist (planner.pyx has no removeCollision)
        raise AttributeError("'PyKinematicPlanner' object has no attribute 'removeCollision'")

    def get_planner_status(self):
        return self.planner.getPlannerStatus().decode("utf-8")
