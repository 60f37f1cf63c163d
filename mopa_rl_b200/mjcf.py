"""MJCF-subset model compiler.

The reference loads its scenes with MuJoCo 2.0's ``mj_loadXML`` twice per worker:
once through mujoco-py for the env (``env/base.py:118-135``) and once inside the
planner (``motion_planners/KinematicPlanner.cpp:62-80``,
``motion_planners/include/mujoco_wrapper.h:75-89``).  MuJoCo is a closed binary that
is not available here, so this module restates the part of the MJCF compiler that
the Sawyer / Pusher scene files under ``env/assets/xml`` actually use:

* ``<include>`` (paths relative to the directory of the top-level file),
* nested ``<default class=...>`` trees, ``class=`` and ``childclass=``,
* ``<compiler angle meshdir inertiafromgeom>``, ``<option>``,
* bodies / inertial / joint / geom / site, ``pos quat euler fromto ref``,
* ``<contact><exclude>``, ``<actuator><position|velocity|motor>``,
* inertia inferred from geoms (primitives analytically, meshes from binary STL
  with the legacy centroid-pyramid rule of MuJoCo <= 2.1).

The output is a :class:`CompiledModel` of flat float64/int32 numpy arrays whose names
follow ``mjModel`` so the rest of the code reads like the reference's callers
(``sim.model.jnt_range`` ...).  Ids follow MuJoCo's ordering rules: bodies in
depth-first document order, joints/geoms/sites body-major.
"""
from __future__ import annotations

import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

# mjtGeom / mjtJoint enums (mujoco.h of MuJoCo 2.0)
GEOM_PLANE, GEOM_HFIELD, GEOM_SPHERE, GEOM_CAPSULE, GEOM_ELLIPSOID, GEOM_CYLINDER, GEOM_BOX, GEOM_MESH = range(8)
JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = range(4)

_GEOM_TYPES = {
    "plane": GEOM_PLANE, "hfield": GEOM_HFIELD, "sphere": GEOM_SPHERE, "capsule": GEOM_CAPSULE,
    "ellipsoid": GEOM_ELLIPSOID, "cylinder": GEOM_CYLINDER, "box": GEOM_BOX, "mesh": GEOM_MESH,
}
_JNT_TYPES = {"free": JNT_FREE, "ball": JNT_BALL, "slide": JNT_SLIDE, "hinge": JNT_HINGE}


# --------------------------------------------------------------------------- math helpers
def _vec(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and len(v) < n and default is not None:
        d = np.array(default, dtype=np.float64)
        d[: len(v)] = v
        v = d
    return v


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def quat_to_mat(q):
    w, x, y, z = q
    return np.array([
        [w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z],
    ])


def mat_to_quat(m):
    # robust branch (Shepperd)
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s])
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = np.array([(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s])
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = np.array([(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s])
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = np.array([(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s])
    return q / np.linalg.norm(q)


def _normalize_quat(q):
    n = np.linalg.norm(q)
    if n < 1e-14:
        return np.array([1.0, 0, 0, 0])
    return q / n


def _axis_angle_quat(axis, angle):
    axis = axis / np.linalg.norm(axis)
    return np.concatenate([[np.cos(angle / 2)], np.sin(angle / 2) * axis])


def _z_to_vec_quat(v):
    """Quaternion rotating the z axis onto unit vector v (MuJoCo's mjuu_z2quat)."""
    v = v / np.linalg.norm(v)
    z = np.array([0.0, 0, 1])
    axis = np.cross(z, v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        return np.array([1.0, 0, 0, 0]) if v[2] > 0 else np.array([0.0, 1, 0, 0])
    ang = np.arctan2(s, v[2])
    return _axis_angle_quat(axis / s, ang)


# --------------------------------------------------------------------------- STL / mesh inertia
def read_stl(path, scale=(1, 1, 1)):
    with open(path, "rb") as f:
        data = f.read()
    ntri = struct.unpack_from("<I", data, 80)[0]
    if 84 + 50 * ntri != len(data):
        # ASCII fallback
        verts = []
        for line in data.decode("ascii", "ignore").splitlines():
            p = line.split()
            if len(p) == 4 and p[0] == "vertex":
                verts.append([float(p[1]), float(p[2]), float(p[3])])
        tri = np.array(verts, dtype=np.float64).reshape(-1, 3, 3)
    else:
        rec = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                            count=ntri, offset=84)
        tri = rec["v"].astype(np.float64)
    return tri * np.asarray(scale, dtype=np.float64)


def mesh_inertia_legacy(tri):
    """Volume, centre of mass and inertia (about the COM, unit density) of a triangle soup.

    MuJoCo <= 2.1 rule: the reference point is the area-weighted centroid of the
    surface; every face forms a pyramid with it and the |volume| of each pyramid is
    used, so non-convex meshes are over-estimated exactly as MuJoCo 2.0 does.
    """
    a, b, c = tri[:, 0], tri[:, 1], tri[:, 2]
    area = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
    cen = ((a + b + c) / 3.0 * area[:, None]).sum(0) / area.sum()
    a, b, c = a - cen, b - cen, c - cen
    vol = np.abs(np.einsum("ij,ij->i", a, np.cross(b, c))) / 6.0
    V = vol.sum()
    com = ((a + b + c) / 4.0 * vol[:, None]).sum(0) / V
    # second moments of each tetra (vertices 0,a,b,c): integral x x^T dV = vol/20 * (sum_i v_i v_i^T + (sum v)(sum v)^T)
    s = a + b + c
    P = np.einsum("i,ijk->jk", vol / 20.0,
                  np.einsum("ij,ik->ijk", a, a) + np.einsum("ij,ik->ijk", b, b) + np.einsum("ij,ik->ijk", c, c)
                  + np.einsum("ij,ik->ijk", s, s))
    P -= V * np.outer(com, com)
    I = np.trace(P) * np.eye(3) - P
    return V, cen + com, I


def mesh_hull_vertices(tri):
    """Vertices of the convex hull of a triangle soup (float32 values, original vertex order).

    MuJoCo collides mesh geoms through the convex hull of their vertices (qhull at compile
    time, libccd support function = arg max of v . dir over the hull vertices).  The hull is
    computed here with the same library family (scipy's qhull); vertices are de-duplicated on
    their float32 values and kept in order of first appearance so that ties of the support
    function resolve identically everywhere (lowest index wins).
    """
    from scipy.spatial import ConvexHull  # host-side scene compilation only

    v = np.asarray(tri, dtype=np.float64).reshape(-1, 3).astype(np.float32)
    _, first = np.unique(v, axis=0, return_index=True)
    v = v[np.sort(first)].astype(np.float64)
    hull = ConvexHull(v)
    if not np.all(hull.equations[:, 3] < 0):
        raise ValueError("mesh origin is not strictly inside its convex hull (needed as the MPR interior point)")
    return v[np.sort(hull.vertices)]


# --------------------------------------------------------------------------- defaults
class _Defaults:
    def __init__(self, parent=None):
        self.attrs = {} if parent is None else {k: dict(v) for k, v in parent.attrs.items()}

    def update(self, tag, attrib):
        self.attrs.setdefault(tag, {}).update(attrib)

    def get(self, tag):
        return self.attrs.get(tag, {})


class CompiledModel:
    """Flat model arrays (``mjModel`` naming).  All float arrays are float64."""

    def __init__(self):
        self.names = {}

    def name2id(self, kind, name):
        return self.names[kind].index(name)

    # mujoco-py style helpers used by the reference env code (env/sawyer/sawyer.py:141-200)
    def body_name2id(self, n):
        return self.name2id("body", n)

    def geom_name2id(self, n):
        return self.name2id("geom", n)

    def site_name2id(self, n):
        return self.name2id("site", n)

    def joint_name2id(self, n):
        return self.name2id("joint", n)

    def get_joint_qpos_addr(self, n):
        j = self.joint_name2id(n)
        a = int(self.jnt_qposadr[j])
        t = int(self.jnt_type[j])
        if t == JNT_FREE:
            return (a, a + 7)
        if t == JNT_BALL:
            return (a, a + 4)
        return a

    def get_joint_qvel_addr(self, n):
        j = self.joint_name2id(n)
        a = int(self.jnt_dofadr[j])
        t = int(self.jnt_type[j])
        if t == JNT_FREE:
            return (a, a + 6)
        if t == JNT_BALL:
            return (a, a + 3)
        return a

    # ---- serialisation (the GPU box has no /root/reference: compiled models travel as .npz)
    _SCALARS = ("nq", "nv", "nu", "nbody", "njnt", "ngeom", "nsite")

    def to_dict(self):
        d = {}
        for k, v in self.__dict__.items():
            if k == "names":
                for kind, lst in v.items():
                    d["names_" + kind] = np.array(lst, dtype=object).astype(str)
            elif isinstance(v, (int, float, str)):
                d[k] = np.array(v)
            else:
                d[k] = np.asarray(v)
        return d

    def save(self, path):
        np.savez_compressed(path, **self.to_dict())

    @classmethod
    def load(cls, path):
        m = cls()
        with np.load(path, allow_pickle=False) as z:
            for k in z.files:
                v = z[k]
                if k.startswith("names_"):
                    m.names[k[6:]] = [str(x) for x in v.tolist()]
                elif v.ndim == 0:
                    v = v.item()
                    setattr(m, k, v)
                else:
                    setattr(m, k, v)
        return m


# --------------------------------------------------------------------------- the compiler
class _Compiler:
    def __init__(self, path):
        self.path = os.path.abspath(path)
        self.basedir = os.path.dirname(self.path)
        self.angle_scale = np.pi / 180.0  # MJCF default angle="degree"
        self.meshdir = ""
        self.inertiafromgeom = "auto"
        self.eulerseq = "xyz"
        self.opt = dict(timestep=0.002, gravity=[0, 0, -9.81], integrator="Euler", cone="pyramidal",
                        iterations=100, noslip_iterations=0, tolerance=1e-8, impratio=1.0, solver="Newton")
        self.meshes = {}
        self.excludes = []
        self.default_root = _Defaults()
        self.classes = {"main": self.default_root}
        self.bodies, self.joints, self.geoms, self.sites, self.actuators = [], [], [], [], []

    # -- xml loading with includes
    def _expand(self, elem):
        out = []
        for child in list(elem):
            if child.tag == "include":
                inc = ET.parse(os.path.join(self.basedir, child.attrib["file"])).getroot()
                self._expand(inc)
                out.extend(list(inc))  # children of <mujocoinclude>/<mujoco> spliced in place
            else:
                self._expand(child)
                out.append(child)
        elem[:] = out

    def _read_defaults(self, elem, parent):
        for child in elem:
            if child.tag == "default":
                name = child.attrib.get("class")
                d = _Defaults(parent)
                if name is not None:
                    self.classes[name] = d
                    self._read_defaults(child, d)
                else:
                    self._read_defaults(child, parent)
            else:
                parent.update(child.tag, child.attrib)
                # children created *before* this update must not see it; MJCF declares
                # element defaults before nested classes in all files we compile.

    def _attrs(self, elem, childclass, tag=None):
        cls = elem.attrib.get("class", childclass)
        d = self.classes[cls] if cls else self.default_root
        a = dict(d.get(tag or elem.tag))
        a.update(elem.attrib)
        return a

    def _orientation(self, a):
        if "quat" in a:
            return _normalize_quat(_vec(a["quat"]))
        if "euler" in a:
            e = _vec(a["euler"]) * self.angle_scale
            q = np.array([1.0, 0, 0, 0])
            for ch, ang in zip(self.eulerseq, e):
                ax = {"x": [1.0, 0, 0], "y": [0, 1.0, 0], "z": [0, 0, 1.0]}[ch.lower()]
                r = _axis_angle_quat(np.array(ax), ang)
                q = quat_mul(q, r) if ch.islower() else quat_mul(r, q)
            return q
        if "axisangle" in a:
            v = _vec(a["axisangle"])
            return _axis_angle_quat(v[:3], v[3] * self.angle_scale)
        if "zaxis" in a:
            return _z_to_vec_quat(_vec(a["zaxis"]))
        return np.array([1.0, 0, 0, 0])

    # -- tree walk
    def _body(self, elem, parent_id, childclass):
        bid = len(self.bodies)
        a = elem.attrib
        childclass = a.get("childclass", childclass)
        b = dict(name=a.get("name", "world" if bid == 0 else f"body{bid}"), parent=parent_id,
                 pos=_vec(a.get("pos"), default=[0, 0, 0]), quat=self._orientation(a),
                 inertial=None, joints=[], geoms=[])
        self.bodies.append(b)
        for child in elem:
            if child.tag == "inertial":
                ia = child.attrib
                if "fullinertia" in ia:
                    raise NotImplementedError("fullinertia")
                b["inertial"] = dict(pos=_vec(ia.get("pos"), default=[0, 0, 0]), quat=self._orientation(ia),
                                     mass=float(ia["mass"]), diag=_vec(ia.get("diaginertia"), default=[0, 0, 0]))
            elif child.tag == "joint" or child.tag == "freejoint":
                ja = self._attrs(child, childclass, "joint")
                jt = JNT_FREE if child.tag == "freejoint" else _JNT_TYPES[ja.get("type", "hinge")]
                rng = _vec(ja.get("range"), default=[0, 0])
                if jt == JNT_HINGE:
                    rng = rng * self.angle_scale
                ref = float(ja.get("ref", 0.0)) * (self.angle_scale if jt == JNT_HINGE else 1.0)
                limited = ja.get("limited", "false") == "true"
                j = dict(name=ja.get("name", f"joint{len(self.joints)}"), type=jt, body=bid,
                         pos=_vec(ja.get("pos"), default=[0, 0, 0]),
                         axis=_vec(ja.get("axis"), default=[0, 0, 1]), limited=limited, range=rng, ref=ref,
                         damping=float(ja.get("damping", 0.0)), armature=float(ja.get("armature", 0.0)),
                         stiffness=float(ja.get("stiffness", 0.0)), margin=float(ja.get("margin", 0.0)),
                         frictionloss=float(ja.get("frictionloss", 0.0)),
                         solref=_vec(ja.get("solreflimit"), default=[0.02, 1.0]),
                         solimp=_vec(ja.get("solimplimit"), 5, default=[0.9, 0.95, 0.001, 0.5, 2.0]))
                if jt in (JNT_SLIDE, JNT_HINGE):
                    j["axis"] = j["axis"] / np.linalg.norm(j["axis"])
                else:
                    j["axis"] = np.array([0.0, 0, 1])
                self.joints.append(j)
                b["joints"].append(len(self.joints) - 1)
            elif child.tag == "geom":
                self._geom(child, bid, childclass)
            elif child.tag == "site":
                sa = self._attrs(child, childclass)
                self.sites.append(dict(name=sa.get("name", f"site{len(self.sites)}"), body=bid,
                                       pos=_vec(sa.get("pos"), default=[0, 0, 0]), quat=self._orientation(sa)))
        for child in elem:
            if child.tag == "body":
                self._body(child, bid, childclass)

    def _geom(self, elem, bid, childclass):
        ga = self._attrs(elem, childclass)
        gt = _GEOM_TYPES[ga.get("type", "sphere")]
        size = _vec(ga.get("size"), 3, default=[0, 0, 0])
        if size is None:
            size = np.zeros(3)
        size = np.concatenate([size, np.zeros(3)])[:3]
        pos = _vec(ga.get("pos"), default=[0, 0, 0])
        quat = self._orientation(ga)
        if "fromto" in ga:
            ft = _vec(ga["fromto"])
            p0, p1 = ft[:3], ft[3:]
            pos = 0.5 * (p0 + p1)
            quat = _z_to_vec_quat(p1 - p0)
            size = np.array([size[0], 0.5 * np.linalg.norm(p1 - p0), 0.0])
        g = dict(name=ga.get("name", ""), type=gt, body=bid, pos=pos, quat=quat, size=size,
                 contype=int(ga.get("contype", 1)), conaffinity=int(ga.get("conaffinity", 1)),
                 condim=int(ga.get("condim", 3)), group=int(ga.get("group", 0)),
                 friction=_vec(ga.get("friction"), 3, default=[1.0, 0.005, 0.0001]),
                 margin=float(ga.get("margin", 0.0)), gap=float(ga.get("gap", 0.0)),
                 solref=_vec(ga.get("solref"), default=[0.02, 1.0]),
                 solimp=_vec(ga.get("solimp"), 5, default=[0.9, 0.95, 0.001, 0.5, 2.0]),
                 solmix=float(ga.get("solmix", 1.0)), density=float(ga.get("density", 1000.0)),
                 mass=(float(ga["mass"]) if "mass" in ga else None), mesh=ga.get("mesh"),
                 rgba=_vec(ga.get("rgba"), default=[0.5, 0.5, 0.5, 1.0]))
        if gt == GEOM_MESH:
            # MuJoCo re-expresses mesh geoms in the mesh's inertial frame; only mass
            # properties of mesh geoms are used by this code base (collision meshes: lift "can").
            m = self.meshes[g["mesh"]]
            g["mesh_vol"], g["mesh_com"], g["mesh_I"] = m["vol"], m["com"], m["I"]
        self.geoms.append(g)
        self.bodies[bid]["geoms"].append(len(self.geoms) - 1)

    # -- geom mass properties (unit: returns mass, com in geom frame, inertia matrix about com in geom frame)
    @staticmethod
    def _geom_inertia(g):
        t, s, rho = g["type"], g["size"], g["density"]
        if t == GEOM_SPHERE:
            V = 4.0 / 3.0 * np.pi * s[0] ** 3
            I = np.eye(3) * (0.4 * s[0] ** 2)
        elif t == GEOM_CAPSULE:
            r, h = s[0], 2 * s[1]
            Vc, Vs = np.pi * r * r * h, 4.0 / 3.0 * np.pi * r ** 3
            V = Vc + Vs
            # MuJoCo: cylinder + sphere split in two hemispheres displaced by h/2
            Ixc = Vc * (r * r / 4 + h * h / 12)
            Izc = Vc * r * r / 2
            Ixs = Vs * (2 * r * r / 5 + h * h / 4 + 3 * r * h / 8)
            Izs = Vs * 2 * r * r / 5
            I = np.diag([Ixc + Ixs, Ixc + Ixs, Izc + Izs]) / V
        elif t == GEOM_CYLINDER:
            r, h = s[0], 2 * s[1]
            V = np.pi * r * r * h
            I = np.diag([(3 * r * r + h * h) / 12, (3 * r * r + h * h) / 12, r * r / 2])
        elif t == GEOM_BOX:
            V = 8 * s[0] * s[1] * s[2]
            I = np.diag([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2]) / 3.0
        elif t == GEOM_ELLIPSOID:
            V = 4.0 / 3.0 * np.pi * s[0] * s[1] * s[2]
            I = np.diag([s[1] ** 2 + s[2] ** 2, s[0] ** 2 + s[2] ** 2, s[0] ** 2 + s[1] ** 2]) / 5.0
        elif t == GEOM_MESH:
            V = g["mesh_vol"]
            mass = g["mass"] if g["mass"] is not None else rho * V
            return mass, g["mesh_com"], g["mesh_I"] / V * mass
        else:
            return 0.0, np.zeros(3), np.zeros((3, 3))
        mass = g["mass"] if g["mass"] is not None else rho * V
        return mass, np.zeros(3), I * mass

    def compile(self):
        root = ET.parse(self.path).getroot()
        self._expand(root)
        for c in root.iter("compiler"):
            if "angle" in c.attrib:
                self.angle_scale = 1.0 if c.attrib["angle"] == "radian" else np.pi / 180.0
            self.meshdir = c.attrib.get("meshdir", self.meshdir)
            self.inertiafromgeom = c.attrib.get("inertiafromgeom", self.inertiafromgeom)
            self.eulerseq = c.attrib.get("eulerseq", self.eulerseq)
        for o in root.iter("option"):
            for k, v in o.attrib.items():
                if k == "gravity":
                    self.opt[k] = [float(x) for x in v.split()]
                elif k in ("integrator", "cone", "solver"):
                    self.opt[k] = v
                elif k in ("iterations", "noslip_iterations"):
                    self.opt[k] = int(v)
                elif k in self.opt:
                    self.opt[k] = float(v)
        for d in root.findall("default"):
            self._read_defaults(d, self.default_root)
        for asset in root.findall("asset"):
            for m in asset.findall("mesh"):
                ma = dict(self.default_root.get("mesh"))
                ma.update(m.attrib)
                f = os.path.join(self.basedir, self.meshdir, ma["file"])
                name = ma.get("name", os.path.splitext(os.path.basename(f))[0])
                tri = read_stl(f, _vec(ma.get("scale"), default=[1, 1, 1]))
                vol, com, I = mesh_inertia_legacy(tri)
                self.meshes[name] = dict(vol=vol, com=com, I=I, ntri=len(tri), file=f, tri=tri, hull=None)
        for c in root.findall("contact"):
            for e in c.findall("exclude"):
                self.excludes.append((e.attrib["body1"], e.attrib["body2"]))
        # world body: all <worldbody> sections merge
        world = ET.Element("body", {"name": "world"})
        for wb in root.findall("worldbody"):
            world.extend(list(wb))
        self._body(world, 0, None)
        self.bodies[0]["parent"] = 0
        for sec in root.findall("actuator"):
            for a in sec:
                aa = self._attrs(a, None)
                act = dict(name=aa.get("name", f"actuator{len(self.actuators)}"), kind=a.tag, joint=aa["joint"],
                           ctrllimited=aa.get("ctrllimited", "false") == "true",
                           ctrlrange=_vec(aa.get("ctrlrange"), default=[0, 0]),
                           forcelimited=aa.get("forcelimited", "false") == "true",
                           forcerange=_vec(aa.get("forcerange"), default=[0, 0]),
                           gear=_vec(aa.get("gear"), 6, default=[1, 0, 0, 0, 0, 0])[0],
                           kp=float(aa.get("kp", 1.0)), kv=float(aa.get("kv", 1.0)))
                self.actuators.append(act)
        return self._finish()

    def _finish(self):
        m = CompiledModel()
        B, J, G, S, A = self.bodies, self.joints, self.geoms, self.sites, self.actuators
        nb = len(B)
        m.nbody, m.njnt, m.ngeom, m.nsite, m.nu = nb, len(J), len(G), len(S), len(A)
        m.names = dict(body=[b["name"] for b in B], joint=[j["name"] for j in J], geom=[g["name"] for g in G],
                       site=[s["name"] for s in S], actuator=[a["name"] for a in A])
        # MuJoCo orders geoms body-major; our walk appends a body's own geoms before descending,
        # and bodies depth-first, which is already body-major in body-id order.
        order = sorted(range(len(G)), key=lambda i: (G[i]["body"], i))
        assert order == list(range(len(G)))
        assert [j["body"] for j in J] == sorted(j["body"] for j in J)

        m.body_parentid = np.array([b["parent"] for b in B], dtype=np.int32)
        m.body_pos = np.array([b["pos"] for b in B])
        m.body_quat = np.array([b["quat"] for b in B])
        m.body_jntnum = np.array([len(b["joints"]) for b in B], dtype=np.int32)
        m.body_jntadr = np.array([b["joints"][0] if b["joints"] else -1 for b in B], dtype=np.int32)
        m.body_geomnum = np.array([len(b["geoms"]) for b in B], dtype=np.int32)
        m.body_geomadr = np.array([b["geoms"][0] if b["geoms"] else -1 for b in B], dtype=np.int32)
        weld = np.zeros(nb, dtype=np.int32)
        for i in range(1, nb):
            weld[i] = i if B[i]["joints"] else weld[B[i]["parent"]]
        m.body_weldid = weld

        # joints / dofs / qpos
        qadr, dadr = 0, 0
        jq, jd, qpos0 = [], [], []
        dof_body, dof_jnt, dof_arm, dof_damp = [], [], [], []
        for ji, j in enumerate(J):
            jq.append(qadr)
            jd.append(dadr)
            b = B[j["body"]]
            if j["type"] == JNT_FREE:
                nqj, nvj = 7, 6
                qpos0.extend(list(b["pos"]) + list(b["quat"]))
            elif j["type"] == JNT_BALL:
                nqj, nvj = 4, 3
                qpos0.extend([1, 0, 0, 0])
            else:
                nqj, nvj = 1, 1
                qpos0.append(j["ref"])
            qadr += nqj
            dadr += nvj
            for _ in range(nvj):
                dof_body.append(j["body"])
                dof_jnt.append(ji)
                dof_arm.append(j["armature"])
                dof_damp.append(j["damping"])
        m.nq, m.nv = qadr, dadr
        m.qpos0 = np.array(qpos0, dtype=np.float64)
        m.jnt_type = np.array([j["type"] for j in J], dtype=np.int32)
        m.jnt_qposadr = np.array(jq, dtype=np.int32)
        m.jnt_dofadr = np.array(jd, dtype=np.int32)
        m.jnt_bodyid = np.array([j["body"] for j in J], dtype=np.int32)
        m.jnt_pos = np.array([j["pos"] for j in J]).reshape(-1, 3)
        m.jnt_axis = np.array([j["axis"] for j in J]).reshape(-1, 3)
        m.jnt_limited = np.array([j["limited"] for j in J], dtype=np.uint8)
        m.jnt_range = np.array([j["range"] for j in J]).reshape(-1, 2)
        m.jnt_ref = np.array([j["ref"] for j in J])
        m.jnt_margin = np.array([j["margin"] for j in J])
        m.jnt_stiffness = np.array([j["stiffness"] for j in J])
        m.jnt_solref = np.array([j["solref"] for j in J]).reshape(-1, 2)
        m.jnt_solimp = np.array([j["solimp"] for j in J]).reshape(-1, 5)
        m.dof_bodyid = np.array(dof_body, dtype=np.int32)
        m.dof_jntid = np.array(dof_jnt, dtype=np.int32)
        m.dof_armature = np.array(dof_arm)
        m.dof_damping = np.array(dof_damp)
        m.body_dofnum = np.array([sum(1 for d in dof_body if d == i) for i in range(nb)], dtype=np.int32)
        m.body_dofadr = np.array([dof_body.index(i) if i in dof_body else -1 for i in range(nb)], dtype=np.int32)
        # dof_parentid: previous dof in the same body, else last dof of the nearest ancestor with dofs
        dpar = []
        for d, bi in enumerate(dof_body):
            if d > 0 and dof_body[d - 1] == bi:
                dpar.append(d - 1)
                continue
            p = B[bi]["parent"]
            while p != 0 and m.body_dofnum[p] == 0:
                p = B[p]["parent"]
            dpar.append(-1 if (p == 0 and m.body_dofnum[0] == 0) else int(m.body_dofadr[p] + m.body_dofnum[p] - 1))
        m.dof_parentid = np.array(dpar, dtype=np.int32)

        # geoms
        m.geom_type = np.array([g["type"] for g in G], dtype=np.int32)
        m.geom_bodyid = np.array([g["body"] for g in G], dtype=np.int32)
        m.geom_pos = np.array([g["pos"] for g in G]).reshape(-1, 3)
        m.geom_quat = np.array([g["quat"] for g in G]).reshape(-1, 4)
        m.geom_size = np.array([g["size"] for g in G]).reshape(-1, 3)
        m.geom_contype = np.array([g["contype"] for g in G], dtype=np.int32)
        m.geom_conaffinity = np.array([g["conaffinity"] for g in G], dtype=np.int32)
        m.geom_condim = np.array([g["condim"] for g in G], dtype=np.int32)
        m.geom_group = np.array([g["group"] for g in G], dtype=np.int32)
        m.geom_friction = np.array([g["friction"] for g in G]).reshape(-1, 3)
        m.geom_margin = np.array([g["margin"] for g in G])
        m.geom_gap = np.array([g["gap"] for g in G])
        m.geom_solref = np.array([g["solref"] for g in G]).reshape(-1, 2)
        m.geom_solimp = np.array([g["solimp"] for g in G]).reshape(-1, 5)
        m.geom_solmix = np.array([g["solmix"] for g in G])
        m.geom_rgba = np.array([g["rgba"] for g in G]).reshape(-1, 4)
        rb = []
        for g in G:
            t, s = g["type"], g["size"]
            if t == GEOM_SPHERE:
                rb.append(s[0])
            elif t == GEOM_CAPSULE:
                rb.append(s[0] + s[1])
            elif t == GEOM_CYLINDER:
                rb.append(np.sqrt(s[0] ** 2 + s[1] ** 2))
            elif t == GEOM_BOX:
                rb.append(np.linalg.norm(s))
            elif t == GEOM_ELLIPSOID:
                rb.append(max(s))
            elif t == GEOM_MESH:
                rb.append(-1.0)  # collidable mesh geoms: filled below from the hull
            else:
                rb.append(0.0)
        m.geom_rbound = np.array(rb)
        # convex hulls of the meshes that collidable geoms reference (lift: "can").  The geom keeps the
        # frame the XML gives it (MuJoCo re-expresses it in the mesh's inertial frame; the world-space
        # hull is the same), the hull vertices are stored in that frame.
        mesh_ids, vert_adr, vert_num, verts = {}, [], [], []
        m.geom_dataid = np.full(len(G), -1, dtype=np.int32)
        for gi, g in enumerate(G):
            if g["type"] != GEOM_MESH or not (g["contype"] or g["conaffinity"]):
                continue
            name = g["mesh"]
            if name not in mesh_ids:
                me = self.meshes[name]
                if me["hull"] is None:
                    me["hull"] = mesh_hull_vertices(me["tri"])
                mesh_ids[name] = len(vert_adr)
                vert_adr.append(sum(vert_num))
                vert_num.append(len(me["hull"]))
                verts.append(me["hull"])
            m.geom_dataid[gi] = mesh_ids[name]
            m.geom_rbound[gi] = np.sqrt((self.meshes[name]["hull"] ** 2).sum(1).max())
        m.nmesh = len(vert_adr)
        m.mesh_vertadr = np.array(vert_adr, dtype=np.int32)
        m.mesh_vertnum = np.array(vert_num, dtype=np.int32)
        m.mesh_vert = np.concatenate(verts).reshape(-1, 3) if verts else np.zeros((0, 3))
        m.collision_mesh_names = np.array(list(mesh_ids.keys())).astype(str)

        m.site_bodyid = np.array([s["body"] for s in S], dtype=np.int32)
        m.site_pos = np.array([s["pos"] for s in S]).reshape(-1, 3)
        m.site_quat = np.array([s["quat"] for s in S]).reshape(-1, 4)

        # body inertial frames
        mass, ipos, iquat, inert = [], [], [], []
        for bi, b in enumerate(B):
            use_geoms = (self.inertiafromgeom == "true") or (self.inertiafromgeom == "auto" and b["inertial"] is None)
            if not use_geoms or bi == 0:
                if b["inertial"] is None or bi == 0:
                    mass.append(0.0), ipos.append(np.zeros(3)), iquat.append(np.array([1.0, 0, 0, 0])), inert.append(np.zeros(3))
                else:
                    I = b["inertial"]
                    mass.append(I["mass"]), ipos.append(I["pos"]), iquat.append(I["quat"]), inert.append(I["diag"])
                continue
            tot, com, parts = 0.0, np.zeros(3), []
            for gi in b["geoms"]:
                g = G[gi]
                mg, cg, Ig = self._geom_inertia(g)
                if mg <= 0:
                    continue
                R = quat_to_mat(g["quat"])
                c = g["pos"] + R @ cg
                parts.append((mg, c, R @ Ig @ R.T))
                tot += mg
                com += mg * c
            if tot <= 0:
                mass.append(0.0), ipos.append(np.zeros(3)), iquat.append(np.array([1.0, 0, 0, 0])), inert.append(np.zeros(3))
                continue
            com /= tot
            I = np.zeros((3, 3))
            for mg, c, Ig in parts:
                d = c - com
                I += Ig + mg * (np.dot(d, d) * np.eye(3) - np.outer(d, d))
            w, V = np.linalg.eigh(I)
            idx = np.argsort(-w)  # MuJoCo sorts principal inertias in decreasing order
            w, V = w[idx], V[:, idx]
            if np.linalg.det(V) < 0:
                V[:, 2] = -V[:, 2]
            mass.append(tot), ipos.append(com), iquat.append(mat_to_quat(V)), inert.append(w)
        m.body_mass = np.array(mass)
        m.body_ipos = np.array(ipos).reshape(-1, 3)
        m.body_iquat = np.array(iquat).reshape(-1, 4)
        m.body_inertia = np.array(inert).reshape(-1, 3)

        # actuators
        m.actuator_trnid = np.array([m.names["joint"].index(a["joint"]) for a in A], dtype=np.int32).reshape(-1)
        m.actuator_kind = np.array([{"motor": 0, "position": 1, "velocity": 2}[a["kind"]] for a in A], dtype=np.int32)
        m.actuator_ctrllimited = np.array([a["ctrllimited"] for a in A], dtype=np.uint8)
        m.actuator_ctrlrange = np.array([a["ctrlrange"] for a in A]).reshape(-1, 2)
        m.actuator_forcelimited = np.array([a["forcelimited"] for a in A], dtype=np.uint8)
        m.actuator_forcerange = np.array([a["forcerange"] for a in A]).reshape(-1, 2)
        m.actuator_gear = np.array([a["gear"] for a in A]).reshape(-1)
        m.actuator_kp = np.array([a["kp"] for a in A]).reshape(-1)
        m.actuator_kv = np.array([a["kv"] for a in A]).reshape(-1)

        m.exclude_body = np.array([[m.names["body"].index(a), m.names["body"].index(b)] for a, b in self.excludes],
                                  dtype=np.int32).reshape(-1, 2)
        o = self.opt
        m.opt_timestep = float(o["timestep"])
        m.opt_gravity = np.array(o["gravity"], dtype=np.float64)
        m.opt_integrator = 1 if o["integrator"] == "RK4" else 0
        m.opt_cone = 1 if o["cone"] == "elliptic" else 0
        m.opt_iterations = int(o["iterations"])
        m.opt_noslip_iterations = int(o["noslip_iterations"])
        m.opt_tolerance = float(o["tolerance"])
        m.opt_impratio = float(o["impratio"])
        m.mesh_names = np.array(list(self.meshes.keys())).astype(str)
        return m


def compile_mjcf(path) -> CompiledModel:
    """Compile an MJCF file (the subset used by the reference assets) into flat arrays."""
    return _Compiler(path).compile()
