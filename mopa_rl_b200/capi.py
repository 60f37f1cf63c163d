"""ctypes binding of libmopa_b200.so (include/mopa_b200.h).

This is the stand-in for the reference's Cython binding (motion_planners/planner.pyx:31-52):
plain pointers and sizes, no torch types.  There is no CPU fallback: if the library is
missing it is built with nvcc, and if no CUDA device is visible planner creation raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build
from .model import make_desc

_LIB = None


class MopaError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if not os.path.exists(path):
            _build.build()
        L = C.CDLL(path)
        _check_abi(L)
        L.mopa_last_error.restype = C.c_char_p
        L.mopa_device_count.restype = C.c_int
        L.mopa_planner_create.restype = C.c_int
        L.mopa_planner_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_double, C.c_double,
                                          C.c_double, C.c_uint64, C.c_int32, C.POINTER(C.c_void_p)]
        L.mopa_planner_destroy.argtypes = [C.c_void_p]
        L.mopa_planner_destroy.restype = None
        L.mopa_planner_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.mopa_planner_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mopa_scene_pair_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_int32]
        L.mopa_is_valid_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]
        L.mopa_is_valid_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.mopa_is_valid_host_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
        L.mopa_is_valid_active_host_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.mopa_planner_set_max_nodes.argtypes = [C.c_void_p, C.c_int32]
        L.mopa_plan_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mopa_plan_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def scene_pair_table(model, ignored_pairs=(), contact_threshold=0.0):
    """Host-only: which canonical candidate pairs the kernels' pair table keeps after the build-time reach analysis.
    Returns (stats dict, kept uint8[n_canonical])."""
    from .model import make_desc
    desc, keep = make_desc(model)
    ign = np.ascontiguousarray(np.array(list(ignored_pairs), dtype=np.int32).reshape(-1, 2))
    st = np.zeros(8, np.int32)
    L = lib()
    check(L.mopa_scene_pair_table(C.byref(desc), _p(ign) if len(ign) else None, len(ign), float(contact_threshold), _p(st), None, 0))
    kept = np.zeros(int(st[4]), np.uint8)
    check(L.mopa_scene_pair_table(C.byref(desc), _p(ign) if len(ign) else None, len(ign), float(contact_threshold), _p(st), _p(kept), len(kept)))
    names = ("entries", "kept", "runs", "dropped", "canonical", "table_bytes", "frame_floats")
    return dict(zip(names, (int(x) for x in st))), kept


def _check_abi(L):
    """A library built from older headers would read mis-laid-out structs (silent memory corruption): compare the struct
    sizes it was compiled with against the ctypes mirrors and refuse to run on a mismatch."""
    if not hasattr(L, "mopa_abi_sizes"):
        raise MopaError("libmopa_b200.so predates the current headers (no mopa_abi_sizes): rebuild with `python -m mopa_rl_b200.build --force`")
    from .dynmodel import DynDesc
    from .envs import EnvBuffers, SawyerTask
    from .model import ModelDesc
    from .rollout import _RolloutConfig

    got = (C.c_int32 * 5)()
    L.mopa_abi_sizes.argtypes = [C.c_void_p]
    if L.mopa_abi_sizes(got) != 0:
        raise MopaError("mopa_abi_sizes failed")
    want = [C.sizeof(t) for t in (ModelDesc, DynDesc, SawyerTask, EnvBuffers, _RolloutConfig)]
    if list(got) != want:
        raise MopaError("libmopa_b200.so is stale: struct sizes %s (library) != %s (bindings); rebuild with "
                        "`python -m mopa_rl_b200.build --force`" % (list(got), want))


def check(rc):
    if rc != 0:
        raise MopaError("libmopa_b200 error %d: %s" % (rc, lib().mopa_last_error().decode("utf-8", "replace")))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


VALID_FAST, VALID_FIRST_PAIR = 0, 1


class NativePlanner:
    """Owner of one ``mopa_planner*`` (the counterpart of the heap ``KinematicPlanner*`` held by
    ``PyKinematicPlanner``, planner.pyx:32-43)."""

    def __init__(self, model, passive_joint_idx, ignored_contacts, contact_threshold, range_, resolution=0.005, seed=0, device=0):
        L = lib()
        self._L = L
        self.model = model
        desc, self._keep = make_desc(model)
        passive = np.ascontiguousarray(np.array(list(passive_joint_idx), dtype=np.int32).reshape(-1))
        ign = np.ascontiguousarray(np.array([tuple(x) for x in ignored_contacts], dtype=np.int32).reshape(-1, 2))
        h = C.c_void_p()
        check(L.mopa_planner_create(C.byref(desc), _p(passive) if len(passive) else None, len(passive),
                                    _p(ign) if len(ign) else None, len(ign), float(contact_threshold), float(range_),
                                    float(resolution), int(seed) & 0xFFFFFFFFFFFFFFFF, int(device), C.byref(h)))
        self.h = h
        nq, npair, nact = C.c_int32(), C.c_int32(), C.c_int32()
        check(L.mopa_planner_info(self.h, C.byref(nq), C.byref(npair), C.byref(nact)))
        self.nq, self.n_pairs, self.n_active = nq.value, npair.value, nact.value
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._L.mopa_planner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pairs(self):
        g1 = np.zeros(self.n_pairs, np.int32)
        g2 = np.zeros(self.n_pairs, np.int32)
        check(self._L.mopa_planner_pairs(self.h, _p(g1), _p(g2)))
        return g1, g2

    def is_valid_host(self, qpos, flags=VALID_FAST, return_words=False):
        q = np.ascontiguousarray(np.atleast_2d(qpos), dtype=np.float64)
        if q.shape[1] != self.nq:
            raise ValueError("state vector has dimension %d but should be nq: %d" % (q.shape[1], self.nq))
        valid = np.zeros(len(q), np.uint8)
        words = np.zeros(len(q), np.uint32)
        check(self._L.mopa_is_valid_host(self.h, _p(q), len(q), _p(valid), _p(words), int(flags)))
        return (valid, words) if return_words else valid

    def is_valid_device(self, qpos_ptr, row_stride, n, result_ptr, flags=VALID_FAST, stream=0):
        """Raw device-pointer entry (ints from torch ``data_ptr()``); enqueues, does not sync."""
        check(self._L.mopa_is_valid_batch(self.h, C.c_void_p(qpos_ptr), int(row_stride), int(n), C.c_void_p(result_ptr),
                                          int(flags), C.c_void_p(stream)))

    def is_valid_host_f32(self, qpos_ptr, row_stride, n, words_ptr, flags=VALID_FAST):
        """Host fp32 rows (ideally pinned) -> result words in host memory; pipelined copies; blocks until done."""
        check(self._L.mopa_is_valid_host_f32(self.h, C.c_void_p(qpos_ptr), int(row_stride), int(n), C.c_void_p(words_ptr), int(flags)))

    def is_valid_active_host_f32(self, active_ptr, n, base_qpos, words_ptr, flags=VALID_FAST):
        """The reference's convention (KinematicPlanner::isValidState): host fp32 states of the n_active planned joints
        (ideally pinned), passive joints from `base_qpos` (nq values) -> result words in host memory; blocks until done."""
        base = np.ascontiguousarray(base_qpos, dtype=np.float32)
        if base.shape != (self.nq,):
            raise ValueError("base_qpos must have dimension nq: %d" % self.nq)
        check(self._L.mopa_is_valid_active_host_f32(self.h, C.c_void_p(active_ptr), int(n), _p(base), C.c_void_p(words_ptr), int(flags)))

    def is_valid_active(self, active, base_qpos, flags=VALID_FAST):
        """numpy convenience over is_valid_active_host_f32: active[n, n_active] -> valid[n] (bool)."""
        a = np.ascontiguousarray(np.atleast_2d(active), dtype=np.float32)
        if a.shape[1] != self.n_active:
            raise ValueError("active states must have dimension n_active: %d" % self.n_active)
        words = np.zeros(len(a), np.uint32)
        self.is_valid_active_host_f32(a.ctypes.data, len(a), base_qpos, words.ctypes.data, flags)
        return (words & 1).astype(bool) if not flags else words

    def set_max_nodes(self, max_nodes):
        check(self._L.mopa_planner_set_max_nodes(self.h, int(max_nodes)))

    def plan_host(self, start, goal, keys, max_iter, max_path=512):
        """n problems from host arrays.  Returns dict(status[n], path_len[n], path[n,max_path,nq], node_ids, iters)."""
        s = np.ascontiguousarray(np.atleast_2d(start), dtype=np.float64)
        g = np.ascontiguousarray(np.atleast_2d(goal), dtype=np.float64)
        if s.shape[1] != self.nq or g.shape != s.shape:
            raise ValueError("start/goal vectors must have dimension nq: %d" % self.nq)
        n = len(s)
        k = np.ascontiguousarray(np.broadcast_to(np.asarray(keys, dtype=np.uint64), (n,)))
        path = np.zeros((n, max_path, self.nq), np.float64)
        ids = np.zeros((n, max_path), np.int32)
        plen = np.zeros(n, np.int32)
        status = np.zeros(n, np.int32)
        iters = np.zeros(n, np.int32)
        check(self._L.mopa_plan_host(self.h, _p(s), _p(g), _p(k), n, int(max_iter), _p(path), _p(ids), int(max_path), _p(plen),
                                     _p(status), _p(iters)))
        return dict(status=status, path_len=plen, path=path, node_ids=ids, iters=iters)

    def plan_device(self, start_ptr, goal_ptr, row_stride, keys_ptr, n, max_iter, path_ptr, ids_ptr, max_path, len_ptr, status_ptr,
                    iters_ptr=0, nodes_ptr=0, stream=0):
        """Raw device-pointer entry; enqueues on `stream`, does not sync."""
        check(self._L.mopa_plan_batch(self.h, C.c_void_p(start_ptr), C.c_void_p(goal_ptr), int(row_stride), C.c_void_p(keys_ptr),
                                      int(n), int(max_iter), C.c_void_p(path_ptr), C.c_void_p(ids_ptr), int(max_path),
                                      C.c_void_p(len_ptr), C.c_void_p(status_ptr), C.c_void_p(iters_ptr or None),
                                      C.c_void_p(nodes_ptr or None), C.c_void_p(stream)))
