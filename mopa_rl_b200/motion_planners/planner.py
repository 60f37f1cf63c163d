"""``PyKinematicPlanner`` — same constructor and methods as the Cython class of the reference
(``motion_planners/planner.pyx:31-52``), implemented over the ctypes C ABI instead of a
C++ ``KinematicPlanner*`` linked against OMPL and MuJoCo.

Behavioural notes (all from ``motion_planners/KinematicPlanner.cpp``):

* ``plan`` returns a list of rows of ``nq`` floats; failure is a single row filled with ``-5``
  (goal state invalid, :181-184) or ``-4`` (no exact solution, :249-250);
* passive joints keep their start values in every row (:166, :238);
* ``goal_bias``, ``simplified_duration`` and ``num_actions`` are accepted and ignored, as in the
  reference (:42-120 never reads them); ``glue_bodies`` must be empty (the Python callers always
  pass ``[]``, ``sampling_based_planner.py:20,41``);
* the wall-clock ``timelimit`` of ``ss->solve(timelimit)`` (:188) becomes an iteration cap:
  ``max_iter = timelimit * ITERS_PER_SECOND`` (override with ``max_iter=``), and sampling uses a
  counter-based generator keyed by ``seed`` and the per-planner call counter, so a plan is a
  pure function of (seed, call index, start, goal).
"""
from __future__ import annotations

import numpy as np

from ..capi import NativePlanner
from ..model import load_model


class PyKinematicPlanner:
    ITERS_PER_SECOND = 1000  # iteration budget standing in for one second of `timelimit`
    MAX_PATH = 1024

    def __init__(self, xml_filename, algo, num_actions, opt, threshold, _range, passive_joint_idx, glue_bodies,
                 ignored_contacts, contact_threshold, goal_bias, is_simplified, simplified_duration, seed, device=0,
                 model=None):
        dec = lambda s: s.decode("utf-8") if isinstance(s, (bytes, bytearray)) else s
        self.xml_filename = dec(xml_filename)
        self.algo = dec(algo)
        self.opt = dec(opt)
        if ".xml" not in self.xml_filename and model is None:
            raise ValueError("XML model file is required")
        if self.algo != "rrt_connect":
            raise NotImplementedError("planner_type %r: only 'rrt_connect' (the reference default, config/motion_planner.py) is built" % self.algo)
        if len(glue_bodies):
            raise NotImplementedError("glue_bodies is never used by the reference's Python callers")
        if is_simplified:
            raise NotImplementedError("is_simplified=True (PathSimplifier) is not built; the reference default is False")
        self.num_actions = num_actions
        self.threshold = float(threshold)
        self._range = float(_range)
        self.passive_joint_idx = [int(i) for i in passive_joint_idx]
        self.ignored_contacts = [(int(a), int(b)) for a, b in ignored_contacts]
        self.contact_threshold = float(contact_threshold)
        self.seed = int(seed)
        self.model = model if model is not None else load_model(self.xml_filename)
        self._native = NativePlanner(self.model, self.passive_joint_idx, self.ignored_contacts, self.contact_threshold,
                                     self._range, 0.005, self.seed, device)
        self._calls = 0
        self.planner_status = b"none"

    # -- reference API ----------------------------------------------------------------------
    def plan(self, start_vec, goal_vec, timelimit, max_iter=None):
        start = np.asarray(start_vec, dtype=np.float64).reshape(-1)
        goal = np.asarray(goal_vec, dtype=np.float64).reshape(-1)
        nq = self.model.nq
        if start.size != nq or goal.size != nq:
            raise ValueError("start/goal vector has dimension %d/%d but should be nq: %d" % (start.size, goal.size, nq))
        if max_iter is None:
            max_iter = max(1, int(round(float(timelimit) * self.ITERS_PER_SECOND)))
        self._calls += 1
        out = self._native.plan_host(start, goal, [self._calls], max_iter, self.MAX_PATH)
        status, n = int(out["status"][0]), int(out["path_len"][0])
        self.last_iters = int(out["iters"][0])
        if status == 0:
            self.planner_status = b"Exact solution"
            self.last_node_ids = out["node_ids"][0, :n].copy()
            return out["path"][0, :n].tolist()
        self.planner_status = b"Invalid goal" if status == -5 else b"Timeout"
        return [[float(status)] * nq]

    def isValidState(self, state_vec):
        return bool(self._native.is_valid_host(np.asarray(state_vec, dtype=np.float64).reshape(1, -1))[0])

    def getPlannerStatus(self):
        return self.planner_status

    # -- batched extensions used by the vectorised rollout driver ------------------------------
    def isValidStates(self, states):
        return self._native.is_valid_host(states).astype(bool)

    @property
    def native(self):
        return self._native
