"""Selects the simulated sub-trees of a compiled scene and packs them as ``mopa_dyn_desc``
(include/mopa_dyn_desc.h) for the env-step kernel and its oracle.

The reference steps the whole ``mjModel`` (``sim.step()``, env/base.py:388-392).  A kinematic
tree can only move if something acts on it, i.e. if it holds an actuated joint or a collidable
geom; the indicator / target "ghost" arms and the target slider hold neither, so they are left
out of the integration and keep the qpos that reset wrote.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .mjcf import JNT_FREE, JNT_HINGE, JNT_SLIDE, quat_mul, quat_to_mat

_I = C.POINTER(C.c_int32)
_D = C.POINTER(C.c_double)


class DynDesc(C.Structure):
    _fields_ = [
        ("nq", C.c_int32), ("nv", C.c_int32), ("nb", C.c_int32), ("nd", C.c_int32), ("nact", C.c_int32),
        ("ngeom", C.c_int32), ("npair", C.c_int32), ("iterations", C.c_int32),
        ("timestep", C.c_double), ("gravity", C.c_double * 3), ("tolerance", C.c_double),
        ("b_parent", _I), ("b_bodyid", _I), ("b_pos", _D), ("b_quat", _D), ("b_rootpos", _D), ("b_rootquat", _D),
        ("b_jtype", _I), ("b_qadr", _I), ("b_vadr", _I), ("b_dadr", _I), ("b_jaxis", _D), ("b_jpos", _D), ("b_qpos0", _D),
        ("b_mass", _D), ("b_ipos", _D), ("b_iquat", _D), ("b_inertia", _D),
        ("d_body", _I), ("d_qadr", _I), ("d_vadr", _I), ("d_armature", _D), ("d_damping", _D), ("d_limited", _I),
        ("d_range", _D), ("d_solref", _D), ("d_solimp", _D), ("d_margin", _D),
        ("a_dof", _I), ("a_kind", _I), ("a_ctrllimited", _I), ("a_forcelimited", _I), ("a_kp", _D), ("a_kv", _D),
        ("a_gear", _D), ("a_ctrlrange", _D), ("a_forcerange", _D),
        ("g_body", _I), ("g_geomid", _I), ("g_type", _I), ("g_pos", _D), ("g_quat", _D), ("g_size", _D), ("g_rbound", _D),
        ("g_margin", _D), ("g_friction", _D), ("g_solref", _D), ("g_solimp", _D), ("g_condim", _I), ("p_g1", _I), ("p_g2", _I),
        ("integrator", C.c_int32), ("pad_", C.c_int32),
    ]


def static_frames(m):
    """World frames (pos, quat) of bodies with no joint in their ancestry (weld id 0)."""
    pos = np.zeros((m.nbody, 3))
    quat = np.tile(np.array([1.0, 0, 0, 0]), (m.nbody, 1))
    for b in range(1, m.nbody):
        if m.body_weldid[b] != 0:
            continue
        p = m.body_parentid[b]
        pos[b] = pos[p] + quat_to_mat(quat[p]) @ m.body_pos[b]
        quat[b] = quat_mul(quat[p], m.body_quat[b])
    return pos, quat


def candidate_pairs(m):
    """Geom pairs that survive MuJoCo's static collision filters."""
    out = []
    excl = {(int(a), int(b)) for a, b in m.exclude_body} | {(int(b), int(a)) for a, b in m.exclude_body}
    for g1 in range(m.ngeom):
        for g2 in range(g1 + 1, m.ngeom):
            b1, b2 = int(m.geom_bodyid[g1]), int(m.geom_bodyid[g2])
            w1, w2 = int(m.body_weldid[b1]), int(m.body_weldid[b2])
            if w1 == w2:
                continue
            if w1 != 0 and w2 != 0:
                wp1, wp2 = int(m.body_weldid[m.body_parentid[w1]]), int(m.body_weldid[m.body_parentid[w2]])
                if wp1 == w2 or wp2 == w1:
                    continue
            if (b1, b2) in excl:
                continue
            if not ((m.geom_contype[g1] & m.geom_conaffinity[g2]) or (m.geom_contype[g2] & m.geom_conaffinity[g1])):
                continue
            if m.geom_type[g1] == 0 and m.geom_type[g2] == 0:
                continue
            out.append((g1, g2))
    return out


class DynModel:
    def __init__(self, m):
        self.model = m
        nb = m.nbody
        # tree root of every moving body = first ancestor-or-self whose parent is static
        root = np.full(nb, -1, dtype=np.int64)
        for b in range(1, nb):
            if m.body_weldid[b] == 0:
                continue
            p = int(m.body_parentid[b])
            root[b] = b if m.body_weldid[p] == 0 else root[p]
        collidable = (m.geom_contype | m.geom_conaffinity) != 0
        acted = set(int(m.jnt_bodyid[j]) for j in m.actuator_trnid)
        live_roots = set()
        for g in range(m.ngeom):
            if collidable[g] and root[m.geom_bodyid[g]] >= 0:
                live_roots.add(int(root[m.geom_bodyid[g]]))
        for b in acted:
            live_roots.add(int(root[b]))
        # A body with several joints (Pusher: the box and the target carry two slide joints each) becomes a chain of
        # simulated bodies, one joint each: massless virtual bodies for all joints but the last, which stays on the
        # body itself.  Virtual entries are listed as -1 - body id so that look-ups by body id find the real one.
        entries = []   # (model body, joint id or -1, first of its chain, last of its chain)
        for b in range(1, nb):
            if root[b] not in live_roots:
                continue
            nj = int(m.body_jntnum[b])
            if nj <= 1:
                entries.append((b, int(m.body_jntadr[b]) if nj == 1 else -1, True, True))
            else:
                js = [int(m.body_jntadr[b]) + k for k in range(nj)]
                if any(int(m.jnt_type[j]) != JNT_SLIDE for j in js):
                    raise NotImplementedError("simulated body %s has several joints that are not all slides" % m.names["body"][b])
                for k, j in enumerate(js):
                    entries.append((b, j, k == 0, k == nj - 1))
        self.bodies = [b if last else -1 - b for b, _, _, last in entries]
        idx = {b: i for i, (b, _, _, last) in enumerate(entries) if last}
        spos, squat = static_frames(m)
        A = lambda *shape: np.zeros(shape)
        n = len(entries)
        b_parent = np.full(n, -1, np.int32)
        b_jtype = np.full(n, -1, np.int32)
        b_qadr, b_vadr, b_dadr = np.full(n, -1, np.int32), np.full(n, -1, np.int32), np.full(n, -1, np.int32)
        b_rootpos, b_rootquat = A(n, 3), np.tile(np.array([1.0, 0, 0, 0]), (n, 1))
        b_jaxis, b_jpos, b_qpos0 = A(n, 3), A(n, 3), A(n)
        b_pos, b_quat = A(n, 3), np.tile(np.array([1.0, 0, 0, 0]), (n, 1))
        b_mass, b_ipos, b_iquat, b_inertia = A(n), A(n, 3), np.tile(np.array([1.0, 0, 0, 0]), (n, 1)), A(n, 3)
        d_body, d_qadr, d_vadr, d_arm, d_damp, d_lim, d_range, d_solref, d_solimp, d_margin = [], [], [], [], [], [], [], [], [], []
        for i, (b, j, first, last) in enumerate(entries):
            p = int(m.body_parentid[b])
            if not first:
                b_parent[i] = i - 1                      # previous joint of the same body; identity offset
            else:
                b_pos[i], b_quat[i] = m.body_pos[b], m.body_quat[b]
                if p in idx:
                    b_parent[i] = idx[p]
                else:
                    b_rootpos[i], b_rootquat[i] = spos[p], squat[p]
            if last:
                b_mass[i], b_ipos[i], b_iquat[i], b_inertia[i] = m.body_mass[b], m.body_ipos[b], m.body_iquat[b], m.body_inertia[b]
            if j >= 0:
                t = int(m.jnt_type[j])
                if t not in (JNT_FREE, JNT_SLIDE, JNT_HINGE):
                    raise NotImplementedError("ball joints")
                if t == JNT_FREE and (p in idx):
                    raise NotImplementedError("free joint below a moving body")
                b_jtype[i], b_qadr[i], b_vadr[i], b_dadr[i] = t, m.jnt_qposadr[j], m.jnt_dofadr[j], len(d_body)
                b_jaxis[i], b_jpos[i], b_qpos0[i] = m.jnt_axis[j], m.jnt_pos[j], m.jnt_ref[j]
                for k in range(6 if t == JNT_FREE else 1):
                    d_body.append(i)
                    d_qadr.append(int(m.jnt_qposadr[j]) + k if (t != JNT_FREE or k < 3) else -1)
                    d_vadr.append(int(m.jnt_dofadr[j]) + k)
                    d_arm.append(m.dof_armature[m.jnt_dofadr[j] + k])
                    d_damp.append(m.dof_damping[m.jnt_dofadr[j] + k])
                    d_lim.append(int(m.jnt_limited[j]) if t != JNT_FREE else 0)
                    d_range.append(m.jnt_range[j])
                    d_solref.append(m.jnt_solref[j])
                    d_solimp.append(m.jnt_solimp[j])
                    d_margin.append(m.jnt_margin[j])
        self.nb, self.nd = n, len(d_body)
        self.dof_vadr = np.array(d_vadr, np.int32)
        self.dof_qadr = np.array(d_qadr, np.int32)
        # actuators whose joint is simulated
        a_dof, a_ids = [], []
        for a in range(m.nu):
            j = int(m.actuator_trnid[a])
            va = int(m.jnt_dofadr[j])
            a_dof.append(d_vadr.index(va))
            a_ids.append(a)
        # contact geoms / pairs
        pairs = [(g1, g2) for g1, g2 in candidate_pairs(m)
                 if (int(m.geom_bodyid[g1]) in idx or int(m.geom_bodyid[g2]) in idx)]
        used = sorted(set(g for pr in pairs for g in pr))
        gidx = {g: i for i, g in enumerate(used)}
        g_body = np.array([idx.get(int(m.geom_bodyid[g]), -1) for g in used], np.int32)
        g_pos, g_quat = A(len(used), 3), A(len(used), 4)
        for i, g in enumerate(used):
            b = int(m.geom_bodyid[g])
            if g_body[i] >= 0:
                g_pos[i], g_quat[i] = m.geom_pos[g], m.geom_quat[g]
            else:
                g_pos[i] = spos[b] + quat_to_mat(squat[b]) @ m.geom_pos[g]
                g_quat[i] = quat_mul(squat[b], m.geom_quat[g])
        self.pairs = pairs
        self.geoms = used
        # Collision meshes in the DYNAMICS: the contact generator works on primitives, so a mesh geom (lift: the can) is
        # replaced by the bounding cylinder of its convex hull about the geom's z axis (can.stl: r = 25.1 mm, half height
        # 40.0 mm, i.e. the can itself up to its bevels).  State validity / planning use the exact hull.
        g_type, g_size, g_rbound = m.geom_type[used].astype(np.int32), np.array(m.geom_size[used]), np.array(m.geom_rbound[used])
        for i, g in enumerate(used):
            if g_type[i] != 7:
                continue
            me = int(m.geom_dataid[g])
            v = m.mesh_vert[m.mesh_vertadr[me]:m.mesh_vertadr[me] + m.mesh_vertnum[me]]
            r, zc, hh = np.sqrt((v[:, :2] ** 2).sum(1).max()), 0.5 * (v[:, 2].max() + v[:, 2].min()), 0.5 * (v[:, 2].max() - v[:, 2].min())
            off = quat_to_mat(m.geom_quat[g]) @ np.array([0.0, 0.0, zc])
            if g_body[i] >= 0:
                g_pos[i] = g_pos[i] + off
            else:
                g_pos[i] = g_pos[i] + quat_to_mat(squat[int(m.geom_bodyid[g])]) @ off
            g_type[i], g_size[i], g_rbound[i] = 5, [r, hh, 0.0], np.sqrt(r * r + hh * hh)
        arr = dict(
            b_parent=b_parent, b_bodyid=np.array(self.bodies, np.int32), b_pos=b_pos, b_quat=b_quat,
            b_rootpos=b_rootpos, b_rootquat=b_rootquat, b_jtype=b_jtype, b_qadr=b_qadr, b_vadr=b_vadr, b_dadr=b_dadr,
            b_jaxis=b_jaxis, b_jpos=b_jpos, b_qpos0=b_qpos0, b_mass=b_mass, b_ipos=b_ipos, b_iquat=b_iquat, b_inertia=b_inertia,
            d_body=np.array(d_body, np.int32), d_qadr=self.dof_qadr, d_vadr=self.dof_vadr, d_armature=np.array(d_arm),
            d_damping=np.array(d_damp), d_limited=np.array(d_lim, np.int32), d_range=np.array(d_range).reshape(-1, 2),
            d_solref=np.array(d_solref).reshape(-1, 2), d_solimp=np.array(d_solimp).reshape(-1, 5), d_margin=np.array(d_margin),
            a_dof=np.array(a_dof, np.int32), a_kind=m.actuator_kind[a_ids].astype(np.int32),
            a_ctrllimited=m.actuator_ctrllimited[a_ids].astype(np.int32), a_forcelimited=m.actuator_forcelimited[a_ids].astype(np.int32),
            a_kp=m.actuator_kp[a_ids], a_kv=m.actuator_kv[a_ids], a_gear=m.actuator_gear[a_ids],
            a_ctrlrange=m.actuator_ctrlrange[a_ids].reshape(-1, 2), a_forcerange=m.actuator_forcerange[a_ids].reshape(-1, 2),
            g_body=g_body, g_geomid=np.array(used, np.int32), g_type=g_type, g_pos=g_pos, g_quat=g_quat,
            g_size=g_size, g_rbound=g_rbound, g_margin=m.geom_margin[used], g_friction=m.geom_friction[used],
            g_solref=m.geom_solref[used], g_solimp=m.geom_solimp[used], g_condim=m.geom_condim[used].astype(np.int32),
            p_g1=np.array([gidx[a] for a, _ in pairs], np.int32), p_g2=np.array([gidx[b] for _, b in pairs], np.int32),
        )
        self.nact = len(a_ids)
        # constraint-row capacity of the env kernel's workspace variant that serves this scene (env_warp.cu: launch_env_warp)
        self.max_rows = 24 if (self.nb <= 14 and len(used) <= 32) else 32
        self._arr = {}
        d = DynDesc()
        d.nq, d.nv, d.nb, d.nd, d.nact = m.nq, m.nv, self.nb, self.nd, self.nact
        d.ngeom, d.npair, d.iterations = len(used), len(pairs), int(m.opt_iterations)
        d.timestep = float(m.opt_timestep)
        d.tolerance = float(m.opt_tolerance)
        d.integrator = int(getattr(m, "opt_integrator", 0))
        for k in range(3):
            d.gravity[k] = float(m.opt_gravity[k])
        for name, ctype in DynDesc._fields_:
            if ctype is _I or ctype is _D:
                a = np.ascontiguousarray(arr[name], dtype=np.int32 if ctype is _I else np.float64)
                if a.size == 0:
                    a = np.zeros(1, a.dtype)
                self._arr[name] = a
                setattr(d, name, a.ctypes.data_as(ctype))
        self.desc = d
