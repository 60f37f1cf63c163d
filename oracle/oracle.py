"""TEST INFRASTRUCTURE — ctypes wrapper around the CPU oracle (oracle/*.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (mopa_rl_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build(force=False):
    """Compile both oracle flavours with the committed Makefile."""
    targets = [os.path.join(_HERE, "libmopa_oracle_f32.so"), os.path.join(_HERE, "libmopa_oracle_f64.so")]
    if force or not all(os.path.exists(t) for t in targets):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return targets


def lib(precision="f32"):
    if precision not in _LIBS:
        path = os.path.join(_HERE, "libmopa_oracle_%s.so" % precision)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_scene_npair.argtypes = [C.c_void_p]
        L.orc_scene_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_is_valid.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_pair_dists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_fk.argtypes = [C.c_void_p] * 8
        L.orc_primitive_dist.restype = C.c_double
        L.orc_primitive_dist.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIBS[precision] = L
    return _LIBS[precision]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleScene:
    """State-validity oracle for one compiled scene (see orc_collide.c)."""

    def __init__(self, model, ignored_pairs=(), contact_threshold=0.0, precision="f32"):
        from mopa_rl_b200.model import make_desc  # data-format helper only (struct layout)

        self.model = model
        self.L = lib(precision)
        self.precision = precision
        desc, self._keep = make_desc(model)
        ign = np.ascontiguousarray(np.array(list(ignored_pairs), dtype=np.int32).reshape(-1, 2))
        self._ign = ign
        self.h = self.L.orc_scene_create(C.byref(desc), _p(ign) if len(ign) else None, len(ign), float(contact_threshold))
        self.npair = self.L.orc_scene_npair(self.h)

    def __del__(self):
        try:
            self.L.orc_scene_destroy(self.h)
        except Exception:
            pass

    def pairs(self):
        g1 = np.zeros(self.npair, np.int32)
        g2 = np.zeros(self.npair, np.int32)
        self.L.orc_scene_pairs(self.h, _p(g1), _p(g2))
        return g1, g2

    def is_valid(self, qpos, with_dist=False):
        q = np.ascontiguousarray(np.atleast_2d(qpos), dtype=np.float64)
        assert q.shape[1] == self.model.nq
        res = np.zeros(len(q), np.uint32)
        md = np.zeros(len(q), np.float64) if with_dist else None
        self.L.orc_is_valid(self.h, _p(q), len(q), _p(res), _p(md) if with_dist else None)
        return (res, md) if with_dist else res

    def pair_dists(self, qpos):
        q = np.ascontiguousarray(qpos, dtype=np.float64)
        d = np.zeros(self.npair, np.float64)
        self.L.orc_pair_dists(self.h, _p(q), _p(d))
        return d

    def fk(self, qpos):
        m = self.model
        q = np.ascontiguousarray(qpos, dtype=np.float64)
        out = dict(body_xpos=np.zeros((m.nbody, 3)), body_xmat=np.zeros((m.nbody, 9)), geom_xpos=np.zeros((m.ngeom, 3)),
                   geom_xmat=np.zeros((m.ngeom, 9)), site_xpos=np.zeros((m.nsite, 3)), site_xmat=np.zeros((m.nsite, 9)))
        self.L.orc_fk(self.h, _p(q), *[_p(out[k]) for k in ("body_xpos", "body_xmat", "geom_xpos", "geom_xmat", "site_xpos", "site_xmat")])
        return out


def primitive_dist(t1, p1, m1, s1, t2, p2, m2, s2, precision="f64"):
    a = [np.ascontiguousarray(x, dtype=np.float64).ravel() for x in (p1, m1, s1, p2, m2, s2)]
    return lib(precision).orc_primitive_dist(int(t1), _p(a[0]), _p(a[1]), _p(a[2]), int(t2), _p(a[3]), _p(a[4]), _p(a[5]))


class OraclePlanner:
    """RRT-Connect oracle for one scene (see orc_plan.c)."""

    def __init__(self, scene: OracleScene, active_qadr, lo, hi, is_so2, range_, resolution=0.005, seed=0, max_nodes=4096):
        self.scene = scene
        L = scene.L
        L.orc_planner_create.restype = C.c_void_p
        L.orc_planner_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                         C.c_double, C.c_uint64, C.c_int]
        L.orc_planner_destroy.argtypes = [C.c_void_p]
        L.orc_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p]
        self.L = L
        a = np.ascontiguousarray(active_qadr, dtype=np.int32)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        so2 = np.ascontiguousarray(is_so2, dtype=np.int32)
        self.h = L.orc_planner_create(scene.h, _p(a), _p(lo), _p(hi), _p(so2), len(a), float(range_), float(resolution),
                                      int(seed) & 0xFFFFFFFFFFFFFFFF, int(max_nodes))
        self.nq = scene.model.nq

    def __del__(self):
        try:
            self.L.orc_planner_destroy(self.h)
        except Exception:
            pass

    def plan(self, start, goal, key, max_iter, max_path=512):
        s = np.ascontiguousarray(start, dtype=np.float64)
        g = np.ascontiguousarray(goal, dtype=np.float64)
        path = np.zeros((max_path, self.nq), np.float64)
        ids = np.zeros(max_path, np.int32)
        n = C.c_int32()
        it = C.c_int32()
        nn = np.zeros(2, np.int32)
        st = self.L.orc_plan(self.h, _p(s), _p(g), int(key) & 0xFFFFFFFFFFFFFFFF, int(max_iter), _p(path), _p(ids), max_path,
                             C.byref(n), C.byref(it), _p(nn))
        return dict(status=st, path=path[: n.value].copy(), node_ids=ids[: n.value].copy(), iters=it.value, n_nodes=nn)


def space_from_model(model, passive_idx):
    """Active joints / bounds as makeCompoundStateSpace derives them (mujoco_ompl_interface.cpp:149-281)."""
    adr, lo, hi, so2 = [], [], [], []
    passive = set(int(i) for i in passive_idx)
    for j in range(model.njnt):
        a = int(model.jnt_qposadr[j])
        if a in passive:
            continue
        adr.append(a)
        if model.jnt_type[j] == 3 and not model.jnt_limited[j]:
            so2.append(1), lo.append(-np.pi), hi.append(np.pi)
        else:
            so2.append(0), lo.append(model.jnt_range[j, 0]), hi.append(model.jnt_range[j, 1])
    return adr, lo, hi, so2


class OracleDyn:
    """Physics-step oracle (orc_dyn.c) for the simulated sub-trees of a scene."""

    def __init__(self, dynmodel, precision="f64"):
        self.dm = dynmodel
        self.L = lib(precision)
        L = self.L
        L.orc_dyn_create.restype = C.c_void_p
        L.orc_dyn_create.argtypes = [C.c_void_p]
        L.orc_dyn_destroy.argtypes = [C.c_void_p]
        L.orc_dyn_enable_contacts.argtypes = [C.c_void_p, C.c_int]
        L.orc_dyn_set_max_rows.argtypes = [C.c_void_p, C.c_int]
        L.orc_dyn_last_contact_force.restype = C.c_double
        L.orc_dyn_last_contacts.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_dyn_step.argtypes = [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3
        L.orc_dyn_forward.argtypes = [C.c_void_p] * 6
        L.orc_dyn_mass_bias.argtypes = [C.c_void_p] * 6
        self.h = L.orc_dyn_create(C.byref(dynmodel.desc))
        if not self.h:
            raise RuntimeError("scene too large for the dynamics oracle")
        self.nd, self.nb = dynmodel.nd, dynmodel.nb
        L.orc_dyn_set_max_rows(self.h, int(getattr(dynmodel, "max_rows", 36)))
        L.orc_dyn_set_integrator.argtypes = [C.c_void_p, C.c_int]
        L.orc_dyn_set_integrator(self.h, int(getattr(dynmodel.model, "opt_integrator", 0)))   # <option integrator="RK4"> (Pusher)

    def __del__(self):
        try:
            self.L.orc_dyn_destroy(self.h)
        except Exception:
            pass

    def enable_contacts(self, on):
        self.L.orc_dyn_enable_contacts(self.h, int(on))

    def forward(self, qpos, qvel):
        q = np.ascontiguousarray(qpos, np.float64)
        v = np.ascontiguousarray(qvel, np.float64)
        bias, xpos, xquat = np.zeros(self.nd), np.zeros((self.nb, 3)), np.zeros((self.nb, 4))
        self.L.orc_dyn_forward(self.h, _p(q), _p(v), _p(bias), _p(xpos), _p(xquat))
        return bias, xpos, xquat

    def step(self, qpos, qvel, ctrl, comp, bias_prev, nsub=1):
        """In-place on copies; returns (qpos, qvel, bias_prev, xpos, xquat, ncon)."""
        q = np.array(qpos, np.float64)
        v = np.array(qvel, np.float64)
        c = np.ascontiguousarray(ctrl, np.float64)
        cm = np.ascontiguousarray(comp, np.int32)
        b = np.array(bias_prev, np.float64)
        xpos, xquat = np.zeros((self.nb, 3)), np.zeros((self.nb, 4))
        ncon = C.c_int32()
        self.L.orc_dyn_step(self.h, _p(q), _p(v), _p(c), _p(cm), _p(b), int(nsub), _p(xpos), _p(xquat), C.byref(ncon))
        self.contact_force = float(self.L.orc_dyn_last_contact_force())   # BaseEnv.get_contact_force() after this step
        g1, g2 = np.zeros(16, np.int32), np.zeros(16, np.int32)
        k = self.L.orc_dyn_last_contacts(_p(g1), _p(g2))
        self.contacts = [(int(self.dm.geoms[g1[i]]), int(self.dm.geoms[g2[i]])) for i in range(k)]   # (geom1, geom2) model ids
        return q, v, b, xpos, xquat, ncon.value

    def mass_bias(self, qpos, qvel):
        q = np.ascontiguousarray(qpos, np.float64)
        v = np.ascontiguousarray(qvel, np.float64)
        M, bias, com = np.zeros((self.nd, self.nd)), np.zeros(self.nd), np.zeros((self.nb, 3))
        self.L.orc_dyn_mass_bias(self.h, _p(q), _p(v), _p(M), _p(bias), _p(com))
        return M, bias, com
