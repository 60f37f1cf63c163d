// Narrow-phase signed distances between convex primitives (device + host).
//
// Replaces the collision functions that MuJoCo 2.0 runs inside mj_fwdPosition when the
// reference's validity checker calls it (motion_planners/src/mujoco_ompl_interface.cpp:932).
// Analytic routines for plane-X, sphere-X, capsule-capsule and box-box (15-axis SAT); a
// Minkowski portal refinement routine for the remaining cylinder / capsule / box pairs
// (MuJoCo 2.0 delegates those to libccd's MPR with tolerance 1e-6, 50 iterations).
// Mesh geoms collide through the convex hull of their vertices (support = extreme hull vertex), always via MPR
// (plane - mesh analytically), like mjc_Convex / mjc_PlaneConvex.
// Geoms are ordered by kind (plane < sphere < capsule < cylinder < box < mesh) exactly as
// mjtGeom orders them, so `a` is always the lower kind.
#pragma once
#include "mopa_math.cuh"

namespace mopa {

enum Kind : int { K_PLANE = 0, K_SPHERE = 1, K_CAPSULE = 2, K_CYLINDER = 3, K_BOX = 4, K_MESH = 5 };
#define MOPA_BIG 1.0e10f

// A geom in world coordinates.  Capsules / cylinders only carry their axis in column 2 of R.
struct Geom {
    V3 c;
    M3 R;
    V3 size;
    int kind;
    const float *hull;   // K_MESH: hull vertices (xyz triplets, geom frame), count in nhull
    int nhull;
};

// index of the hull vertex that is extreme along the LOCAL direction l (first maximum)
static MOPA_HD_COLD int hull_extreme(const float *hull, int n, const V3 &l) {
    int best = 0;
    float bd = dot(V3{hull[0], hull[1], hull[2]}, l);
    for (int i = 1; i < n; i++) {
        float di = dot(V3{hull[3 * i], hull[3 * i + 1], hull[3 * i + 2]}, l);
        if (di > bd) { bd = di; best = i; }
    }
    return best;
}
// `thr`: distances above it are never looked at by the validity predicate, so when the hull's bounding sphere (radius in
// g.size.z, inflated by 1e-6 against rounding) clears the plane by more than thr the 100-odd vertices are not scanned.
MOPA_HD float plane_mesh(const Geom &p, const Geom &g, float thr) {
    V3 n = col(p.R, 2), d = g.c - p.c;
    if (dot(n, d) - g.size.z * 1.000001f > thr) return MOPA_BIG;
    V3 l = mulMTV(g.R, n);
    const float *v = g.hull + 3 * hull_extreme(g.hull, g.nhull, neg(l));
    return dot(n, d) + dot(V3{v[0], v[1], v[2]}, l);
}

MOPA_HD float plane_sphere(const Geom &p, const Geom &s) { return dot(col(p.R, 2), s.c - p.c) - s.size.x; }
MOPA_HD float plane_capsule(const Geom &p, const Geom &g) {
    V3 n = col(p.R, 2), a = col(g.R, 2), d = g.c - p.c;
    float hc = dot(n, d), ha = dot(n, a) * g.size.y;
    return (hc - fabsf(ha)) - g.size.x;
}
MOPA_HD float plane_cylinder(const Geom &p, const Geom &g) {
    V3 n = col(p.R, 2), a = col(g.R, 2), d = g.c - p.c;
    float hc = dot(n, d), na = dot(n, a);
    float s2 = fmaxf(0.0f, fmaf(-na, na, 1.0f));
    return (hc - fabsf(na) * g.size.y) - g.size.x * sqrtf(s2);
}
MOPA_HD float plane_box(const Geom &p, const Geom &g) {
    V3 n = col(p.R, 2), d = g.c - p.c;
    V3 l = mulMTV(g.R, n);
    float ext = fmaf(fabsf(l.z), g.size.z, fmaf(fabsf(l.y), g.size.y, fabsf(l.x) * g.size.x));
    return dot(n, d) - ext;
}
MOPA_HD float sphere_sphere(const Geom &a, const Geom &b) { return (len(b.c - a.c) - a.size.x) - b.size.x; }
MOPA_HD float point_seg(const V3 &p, const V3 &c, const V3 &a, float h) {
    V3 d = p - c;
    float t = fminf(h, fmaxf(-h, dot(d, a)));
    return len(madd(d, -t, a));
}
MOPA_HD float sphere_capsule(const Geom &s, const Geom &g) {
    return (point_seg(s.c, g.c, col(g.R, 2), g.size.y) - s.size.x) - g.size.x;
}
MOPA_HD float capsule_capsule(const Geom &g1, const Geom &g2) {
    V3 a1 = col(g1.R, 2), a2 = col(g2.R, 2);
    float h1 = g1.size.y, h2 = g2.size.y;
    V3 r = g1.c - g2.c;
    float b = dot(a1, a2), c = dot(a1, r), f = dot(a2, r);
    float den = fmaf(-b, b, 1.0f);
    float s, t;
    if (den > 1e-6f) {
        s = fmaf(b, f, -c) / den;
        s = fminf(h1, fmaxf(-h1, s));
    } else
        s = 0.0f;
    t = fmaf(b, s, f);
    if (t < -h2) {
        t = -h2;
        s = fminf(h1, fmaxf(-h1, fmaf(b, t, -c)));
    } else if (t > h2) {
        t = h2;
        s = fminf(h1, fmaxf(-h1, fmaf(b, t, -c)));
    }
    V3 w{fmaf(-t, a2.x, fmaf(s, a1.x, r.x)), fmaf(-t, a2.y, fmaf(s, a1.y, r.y)), fmaf(-t, a2.z, fmaf(s, a1.z, r.z))};
    return (len(w) - g1.size.x) - g2.size.x;
}
MOPA_HD float sphere_cylinder(const Geom &s, const Geom &g) {
    V3 a = col(g.R, 2), d = s.c - g.c;
    float z = dot(d, a);
    V3 w = madd(d, -z, a);
    float dr = len(w) - g.size.x, dz = fabsf(z) - g.size.y;
    float core;
    if (dr <= 0 && dz <= 0) core = fmaxf(dr, dz);
    else if (dz <= 0) core = dr;
    else if (dr <= 0) core = dz;
    else core = sqrtf(fmaf(dr, dr, dz * dz));
    return core - s.size.x;
}
MOPA_HD float sphere_box(const Geom &s, const Geom &g) {
    V3 p = mulMTV(g.R, s.c - g.c);
    V3 e{fabsf(p.x) - g.size.x, fabsf(p.y) - g.size.y, fabsf(p.z) - g.size.z};
    float core;
    if (e.x <= 0 && e.y <= 0 && e.z <= 0) core = fmaxf(e.x, fmaxf(e.y, e.z));
    else {
        V3 o{fmaxf(e.x, 0.0f), fmaxf(e.y, 0.0f), fmaxf(e.z, 0.0f)};
        core = len(o);
    }
    return core - s.size.x;
}
// max over the 15 separating axes of the signed separation (<0: minus penetration depth)
MOPA_HD float box_box(const Geom &g1, const Geom &g2) {
    const float *m1 = g1.R.m, *m2 = g2.R.m;
    const float sz1[3] = {g1.size.x, g1.size.y, g1.size.z}, sz2[3] = {g2.size.x, g2.size.y, g2.size.z};
    V3 Tv = mulMTV(g1.R, g2.c - g1.c);
    const float T[3] = {Tv.x, Tv.y, Tv.z};
    float Rm[9], A[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            Rm[3 * i + j] = fmaf(m1[6 + i], m2[6 + j], fmaf(m1[3 + i], m2[3 + j], m1[i] * m2[j]));
            A[3 * i + j] = fabsf(Rm[3 * i + j]);
        }
    float best = -MOPA_BIG;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float rb = fmaf(A[3 * i + 2], sz2[2], fmaf(A[3 * i + 1], sz2[1], A[3 * i] * sz2[0]));
        best = fmaxf(best, (fabsf(T[i]) - sz1[i]) - rb);
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
        float ra = fmaf(A[6 + j], sz1[2], fmaf(A[3 + j], sz1[1], A[j] * sz1[0]));
        float tp = fmaf(T[2], Rm[6 + j], fmaf(T[1], Rm[3 + j], T[0] * Rm[j]));
        best = fmaxf(best, (fabsf(tp) - ra) - sz2[j]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            float l2 = fmaf(-Rm[3 * i + j], Rm[3 * i + j], 1.0f);
            if (l2 < 1e-6f) continue;
            float ra = fmaf(sz1[i2], A[3 * i1 + j], sz1[i1] * A[3 * i2 + j]);
            float rb = fmaf(sz2[j2], A[3 * i + j1], sz2[j1] * A[3 * i + j2]);
            float tp = fmaf(T[i2], Rm[3 * i1 + j], -(T[i1] * Rm[3 * i2 + j]));
            best = fmaxf(best, ((fabsf(tp) - ra) - rb) / sqrtf(l2));
        }
    }
    return best;
}

// ------------------------------------------------------------------ Minkowski portal refinement
// MESH = false drops the hull / sphere branches at compile time: scenes without mesh colliders (push, assembly) run
// exactly the code (and register budget) they ran before meshes existed.
template <bool MESH>
MOPA_HD V3 support(const Geom &g, const V3 &dir) {
    if (MESH && g.kind == K_MESH) {
        V3 l = mulMTV(g.R, dir);
        const float *v = g.hull + 3 * hull_extreme(g.hull, g.nhull, l);
        return g.c + mulMV(g.R, V3{v[0], v[1], v[2]});
    }
    if (MESH && g.kind == K_SPHERE) return V3{fmaf(dir.x, g.size.x, g.c.x), fmaf(dir.y, g.size.x, g.c.y), fmaf(dir.z, g.size.x, g.c.z)};
    if (g.kind == K_BOX) {
        V3 l = mulMTV(g.R, dir);
        V3 p{l.x >= 0 ? g.size.x : -g.size.x, l.y >= 0 ? g.size.y : -g.size.y, l.z >= 0 ? g.size.z : -g.size.z};
        return g.c + mulMV(g.R, p);
    }
    V3 a = col(g.R, 2);
    float z = dot(dir, a);
    float hs = z >= 0 ? g.size.y : -g.size.y;
    V3 base = madd(g.c, hs, a);
    if (g.kind == K_CYLINDER) {
        V3 w = madd(dir, -z, a);
        float n = len(w);
        float k = n > 1e-12f ? g.size.x / n : 0.0f;
        return V3{fmaf(w.x, k, base.x), fmaf(w.y, k, base.y), fmaf(w.z, k, base.z)};
    }
    return V3{fmaf(dir.x, g.size.x, base.x), fmaf(dir.y, g.size.x, base.y), fmaf(dir.z, g.size.x, base.z)};
}
template <bool MESH>
MOPA_HD V3 msupport(const Geom &g1, const Geom &g2, const V3 &dir) { return support<MESH>(g1, dir) - support<MESH>(g2, neg(dir)); }
MOPA_HD void normalize(V3 &v) {
    float n = len(v);
    if (n < 1e-30f) return;
    v.x /= n; v.y /= n; v.z /= n;
}
#define MPR_EPS 1.1920929e-07f
#define MPR_TOL 1e-6f
#define MPR_MAXIT 50
MOPA_HD bool is_zero(float x) { return fabsf(x) < MPR_EPS; }

MOPA_HD float origin_tri_dist2(const V3 &a, const V3 &b, const V3 &c) {
    V3 ab = b - a, ac = c - a, ap = neg(a);
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) return dot(a, a);
    V3 bp = neg(b);
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) return dot(b, b);
    float vc = fmaf(d1, d4, -(d3 * d2));
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        float v = d1 / (d1 - d3);
        V3 q = madd(a, v, ab);
        return dot(q, q);
    }
    V3 cp = neg(c);
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) return dot(c, c);
    float vb = fmaf(d5, d2, -(d1 * d6));
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        float w = d2 / (d2 - d6);
        V3 q = madd(a, w, ac);
        return dot(q, q);
    }
    float va = fmaf(d3, d6, -(d5 * d4));
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        V3 q = madd(b, w, c - b);
        return dot(q, q);
    }
    V3 n = cross(ab, ac);
    float nn = dot(n, n);
    if (nn < 1e-30f) return dot(a, a);
    float k = dot(n, a);
    return (k * k) / nn;
}

// true + depth when the two convex geoms intersect
template <bool MESH>
MOPA_HD bool mpr_penetration(const Geom &g1, const Geom &g2, float *depth) {
    V3 v0 = g1.c - g2.c, v1, v2, v3, v4, dir, va;
    float d;
    if (v0.x == 0 && v0.y == 0 && v0.z == 0) v0.x = MPR_EPS * 10.0f;
    dir = neg(v0);
    normalize(dir);
    v1 = msupport<MESH>(g1, g2, dir);
    d = dot(v1, dir);
    if (is_zero(d) || d < 0) return false;
    dir = cross(v0, v1);
    if (is_zero(dot(dir, dir))) {
        if (v1.x == 0 && v1.y == 0 && v1.z == 0) { *depth = 0; return true; }
        *depth = len(v1);
        return true;
    }
    normalize(dir);
    v2 = msupport<MESH>(g1, g2, dir);
    d = dot(v2, dir);
    if (is_zero(d) || d < 0) return false;
    dir = cross(v1 - v0, v2 - v0);
    normalize(dir);
    d = dot(dir, v0);
    if (d > 0) {
        V3 t = v1; v1 = v2; v2 = t;
        dir = neg(dir);
    }
    int it = 0;
    for (;;) {
        if (++it > MPR_MAXIT) return false;
        v3 = msupport<MESH>(g1, g2, dir);
        d = dot(v3, dir);
        if (is_zero(d) || d < 0) return false;
        bool cont = false;
        va = cross(v1, v3);
        d = dot(va, v0);
        if (d < 0 && !is_zero(d)) { v2 = v3; cont = true; }
        if (!cont) {
            va = cross(v3, v2);
            d = dot(va, v0);
            if (d < 0 && !is_zero(d)) { v1 = v3; cont = true; }
        }
        if (!cont) break;
        dir = cross(v1 - v0, v2 - v0);
        normalize(dir);
    }
    bool inside = false;
    for (it = 0;; it++) {
        dir = cross(v2 - v1, v3 - v1);
        normalize(dir);
        if (!inside) {
            d = dot(dir, v1);
            if (is_zero(d) || d > 0) inside = true;
        }
        v4 = msupport<MESH>(g1, g2, dir);
        float dv4 = dot(v4, dir);
        float dmin = fminf(dv4 - dot(v1, dir), fminf(dv4 - dot(v2, dir), dv4 - dot(v3, dir)));
        bool reached = (dmin <= MPR_TOL);
        if (!inside) {
            if (!(is_zero(dv4) || dv4 > 0) || reached || it >= MPR_MAXIT) return false;
        } else if (reached || it >= MPR_MAXIT) {
            *depth = sqrtf(origin_tri_dist2(v1, v2, v3));
            return true;
        }
        va = cross(v4, v0);
        d = dot(v1, va);
        if (d > 0) {
            d = dot(v2, va);
            if (d > 0) v1 = v4; else v3 = v4;
        } else {
            d = dot(v3, va);
            if (d > 0) v2 = v4; else v1 = v4;
        }
    }
}

// The same routine as a resumable state machine: one Minkowski support evaluation per trip, then the bookkeeping of the
// phase the item is in.  Every arithmetic expression is the one mpr_penetration evaluates, in the same order, so the
// verdict and the depth are bit-identical; only the control flow is unrolled so that the lanes of a warp can work on
// different items that are at different iterations (the iteration count has a heavy tail: most items finish within a
// few trips, a few run to MPR_MAXIT) and pick up a new item as soon as theirs is finished.
enum { MPR_RUNNING = 0, MPR_SEPARATE = 1, MPR_PENETRATING = 2 };
struct MprSM {
    V3 v0, v1, v2, v3, dir;
    int phase, it;   // phase 0..3: the support point being computed is v1 / v2 / v3 (portal discovery) / v4 (refinement)
    bool inside;
};
MOPA_HD void mpr_begin(MprSM &m, const Geom &g1, const Geom &g2) {
    m.v0 = g1.c - g2.c;
    if (m.v0.x == 0 && m.v0.y == 0 && m.v0.z == 0) m.v0.x = MPR_EPS * 10.0f;
    m.dir = neg(m.v0);
    normalize(m.dir);
    m.phase = 0; m.it = 0; m.inside = false;
    m.v1 = m.v2 = m.v3 = V3{0, 0, 0};
}
MOPA_HD void mpr_refine_dir(MprSM &m) {   // head of a refinement iteration: portal normal, inside test
    m.dir = cross(m.v2 - m.v1, m.v3 - m.v1);
    normalize(m.dir);
    if (!m.inside) {
        const float d = dot(m.dir, m.v1);
        if (is_zero(d) || d > 0) m.inside = true;
    }
}
template <bool MESH>
MOPA_HD int mpr_trip(MprSM &m, const Geom &g1, const Geom &g2, float *depth) {
    const V3 s = msupport<MESH>(g1, g2, m.dir);
    float d;
    if (m.phase == 0) {
        m.v1 = s;
        d = dot(m.v1, m.dir);
        if (is_zero(d) || d < 0) return MPR_SEPARATE;
        m.dir = cross(m.v0, m.v1);
        if (is_zero(dot(m.dir, m.dir))) {
            *depth = (m.v1.x == 0 && m.v1.y == 0 && m.v1.z == 0) ? 0.0f : len(m.v1);
            return MPR_PENETRATING;
        }
        normalize(m.dir);
        m.phase = 1;
        return MPR_RUNNING;
    }
    if (m.phase == 1) {
        m.v2 = s;
        d = dot(m.v2, m.dir);
        if (is_zero(d) || d < 0) return MPR_SEPARATE;
        m.dir = cross(m.v1 - m.v0, m.v2 - m.v0);
        normalize(m.dir);
        d = dot(m.dir, m.v0);
        if (d > 0) {
            const V3 t = m.v1; m.v1 = m.v2; m.v2 = t;
            m.dir = neg(m.dir);
        }
        m.it = 0;
        m.phase = 2;
        return MPR_RUNNING;
    }
    if (m.phase == 2) {
        if (++m.it > MPR_MAXIT) return MPR_SEPARATE;
        m.v3 = s;
        d = dot(m.v3, m.dir);
        if (is_zero(d) || d < 0) return MPR_SEPARATE;
        bool cont = false;
        V3 va = cross(m.v1, m.v3);
        d = dot(va, m.v0);
        if (d < 0 && !is_zero(d)) { m.v2 = m.v3; cont = true; }
        if (!cont) {
            va = cross(m.v3, m.v2);
            d = dot(va, m.v0);
            if (d < 0 && !is_zero(d)) { m.v1 = m.v3; cont = true; }
        }
        if (cont) {
            m.dir = cross(m.v1 - m.v0, m.v2 - m.v0);
            normalize(m.dir);
            return MPR_RUNNING;
        }
        m.inside = false;
        m.it = 0;
        m.phase = 3;
        mpr_refine_dir(m);
        return MPR_RUNNING;
    }
    const V3 v4 = s;
    const float dv4 = dot(v4, m.dir);
    const float dmin = fminf(dv4 - dot(m.v1, m.dir), fminf(dv4 - dot(m.v2, m.dir), dv4 - dot(m.v3, m.dir)));
    const bool reached = (dmin <= MPR_TOL);
    if (!m.inside) {
        if (!(is_zero(dv4) || dv4 > 0) || reached || m.it >= MPR_MAXIT) return MPR_SEPARATE;
    } else if (reached || m.it >= MPR_MAXIT) {
        *depth = sqrtf(origin_tri_dist2(m.v1, m.v2, m.v3));
        return MPR_PENETRATING;
    }
    const V3 va = cross(v4, m.v0);
    d = dot(m.v1, va);
    if (d > 0) {
        d = dot(m.v2, va);
        if (d > 0) m.v1 = v4; else m.v3 = v4;
    } else {
        d = dot(m.v3, va);
        if (d > 0) m.v2 = v4; else m.v1 = v4;
    }
    m.it++;
    mpr_refine_dir(m);
    return MPR_RUNNING;
}

// Conservative pre-test for the pairs that go to MPR (capsule / cylinder against capsule / cylinder / box): true when
// the shapes are certainly more than 1e-4 apart, in which case MPR cannot report a penetration at all (the validity
// predicate needs dist <= contact_threshold <= 0).  A cylinder lies inside the capsule with the same axis, radius and
// half length, so segment distances minus the radii bound the true distance from below.  This never changes a result,
// it only skips the expensive, badly diverging portal refinement for the many near-but-separate pairs.
MOPA_HD bool mpr_certainly_separate(const Geom &a, const Geom &b) {
    const float eps = 1e-4f;
    if (b.kind == K_MESH) return false;   // hulls: bounding spheres only
    if (b.kind == K_BOX) {   // a: capsule / cylinder.  Per box axis: gap between the slab and the projected segment
        const V3 ax = col(a.R, 2), d = a.c - b.c;
        const V3 l = mulMTV(b.R, d), al = mulMTV(b.R, ax);
        const float gx = fmaxf(0.0f, (fabsf(l.x) - fabsf(al.x) * a.size.y) - b.size.x);
        const float gy = fmaxf(0.0f, (fabsf(l.y) - fabsf(al.y) * a.size.y) - b.size.y);
        const float gz = fmaxf(0.0f, (fabsf(l.z) - fabsf(al.z) * a.size.y) - b.size.z);
        const float r = a.size.x + eps;
        return fmaf(gz, gz, fmaf(gy, gy, gx * gx)) > r * r;
    }
    if (a.kind == K_BOX) return false;
    // segment - segment distance (both are capsules / cylinders)
    const V3 a1 = col(a.R, 2), a2 = col(b.R, 2), r = a.c - b.c;
    const float h1 = a.size.y, h2 = b.size.y;
    const float bb = dot(a1, a2), c = dot(a1, r), f = dot(a2, r), den = fmaf(-bb, bb, 1.0f);
    float s = den > 1e-6f ? fminf(h1, fmaxf(-h1, fmaf(bb, f, -c) / den)) : 0.0f;
    float t = fmaf(bb, s, f);
    if (t < -h2) { t = -h2; s = fminf(h1, fmaxf(-h1, fmaf(bb, t, -c))); }
    else if (t > h2) { t = h2; s = fminf(h1, fmaxf(-h1, fmaf(bb, t, -c))); }
    const V3 w{fmaf(-t, a2.x, fmaf(s, a1.x, r.x)), fmaf(-t, a2.y, fmaf(s, a1.y, r.y)), fmaf(-t, a2.z, fmaf(s, a1.z, r.z))};
    const float rr = a.size.x + b.size.x + eps;
    return dot(w, w) > rr * rr;
}

// dispatch classes (pair of kinds, a.kind <= b.kind)
enum PairClass : int {
    PC_PLANE_SPHERE = 0, PC_PLANE_CAPSULE, PC_PLANE_CYLINDER, PC_PLANE_BOX, PC_SPHERE_SPHERE, PC_SPHERE_CAPSULE,
    PC_SPHERE_CYLINDER, PC_SPHERE_BOX, PC_CAPSULE_CAPSULE, PC_PLANE_MESH, PC_BOX_BOX, PC_MPR, PC_NONE
};
MOPA_HD int pair_class(int ka, int kb) {
    if (ka == K_PLANE) return kb == K_SPHERE ? PC_PLANE_SPHERE : kb == K_CAPSULE ? PC_PLANE_CAPSULE : kb == K_CYLINDER ? PC_PLANE_CYLINDER : kb == K_BOX ? PC_PLANE_BOX : kb == K_MESH ? PC_PLANE_MESH : PC_NONE;
    if (ka == K_SPHERE && kb != K_MESH) return kb == K_SPHERE ? PC_SPHERE_SPHERE : kb == K_CAPSULE ? PC_SPHERE_CAPSULE : kb == K_CYLINDER ? PC_SPHERE_CYLINDER : PC_SPHERE_BOX;
    if (ka == K_CAPSULE && kb == K_CAPSULE) return PC_CAPSULE_CAPSULE;
    if (ka == K_BOX && kb == K_BOX) return PC_BOX_BOX;
    return PC_MPR;
}
// cheap classes, evaluated inline by the owning thread
template <bool MESH>
MOPA_HD float cheap_dist(int cls, const Geom &a, const Geom &b, float thr) {
    switch (cls) {
    case PC_PLANE_SPHERE: return plane_sphere(a, b);
    case PC_PLANE_CAPSULE: return plane_capsule(a, b);
    case PC_PLANE_CYLINDER: return plane_cylinder(a, b);
    case PC_PLANE_BOX: return plane_box(a, b);
    case PC_SPHERE_SPHERE: return sphere_sphere(a, b);
    case PC_SPHERE_CAPSULE: return sphere_capsule(a, b);
    case PC_SPHERE_CYLINDER: return sphere_cylinder(a, b);
    case PC_SPHERE_BOX: return sphere_box(a, b);
    case PC_CAPSULE_CAPSULE: return capsule_capsule(a, b);
    case PC_PLANE_MESH: return MESH ? plane_mesh(a, b, thr) : MOPA_BIG;
    default: return MOPA_BIG;
    }
}

}  // namespace mopa
