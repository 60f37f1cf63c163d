/* TEST INFRASTRUCTURE — CPU oracle for the MoPA-RL hot path.  Not part of the product.
 *
 * State-validity oracle: a plain sequential restatement of
 *   MujocoStateValidityChecker::isValid   motion_planners/src/mujoco_ompl_interface.cpp:909-978
 *   KinematicPlanner::isValidState        motion_planners/KinematicPlanner.cpp:253-286
 * i.e. write qpos, run forward kinematics + collision detection, and declare the state
 * invalid iff some contact whose ordered geom pair is not in `ignored_contacts` has
 * dist <= contact_threshold.
 *
 * The arithmetic lives in MuJoCo 2.0 (closed binary, absent here; README.md:19-28 of the
 * reference) -> PARITY UNPINNED against MuJoCo itself.  What is restated:
 *   B.1  mj_kinematics: bodies in id order, hinge / slide / free joints, geom frames
 *   B.3  candidate-pair filters: same weld body, weld parent-child (unless world),
 *        <exclude>, contype/conaffinity, bounding spheres + margin
 *   B.5  signed distance per pair.  Analytic: plane-X, sphere-{sphere,capsule,cylinder,box},
 *        capsule-capsule, box-box (15-axis SAT).  Every other pair goes through a
 *        Minkowski-portal-refinement routine (the algorithm of libccd's ccdMPRPenetration,
 *        which MuJoCo 2.0 uses for all cylinder pairs; tolerance 1e-6, <= 50 iterations).
 * Every geom is evaluated from the raw model arrays on every query: no pre-computation,
 * no culling beyond MuJoCo's own bounding-sphere test.
 */
#include <stdlib.h>
#include <string.h>

#include "../include/mopa_model_desc.h"
#include "orc_math.h"

typedef struct {
    int nq, nbody, njnt, ngeom, nsite;
    int *body_parentid, *body_weldid, *body_jntadr, *body_jntnum;
    R *body_pos, *body_quat;
    int *jnt_type, *jnt_qposadr, *jnt_bodyid, *jnt_limited;
    R *jnt_pos, *jnt_axis, *jnt_range, *qpos0;
    int *geom_type, *geom_bodyid, *geom_contype, *geom_conaffinity;
    R *geom_pos, *geom_mat, *geom_size, *geom_margin, *geom_rbound;
    int *geom_dataid, *mesh_vertadr, *mesh_vertnum; /* convex-hull vertices of collision meshes */
    R *mesh_vert;
    int *site_bodyid;
    R *site_pos, *site_mat;
    int npair;
    int *pair_g1, *pair_g2;
    R threshold;
    /* scratch (one oracle instance = one thread) */
    R *xpos, *xquat, *xmat, *gpos, *gmat;
} orc_scene;

static int *dupi(const int32_t *p, int n) {
    int *r = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) r[i] = p[i];
    return r;
}
static R *dupr(const double *p, int n) {
    R *r = (R *)malloc(sizeof(R) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) r[i] = (R)p[i];
    return r;
}

/* ---------------------------------------------------------------- candidate pairs (B.3) */
static int pair_allowed(const orc_scene *s, const mopa_model_desc *d, int g1, int g2) {
    int b1 = s->geom_bodyid[g1], b2 = s->geom_bodyid[g2];
    int w1 = s->body_weldid[b1], w2 = s->body_weldid[b2];
    if (w1 == w2) return 0;
    if (w1 != 0 && w2 != 0) {
        int wp1 = s->body_weldid[s->body_parentid[w1]], wp2 = s->body_weldid[s->body_parentid[w2]];
        if (wp1 == w2 || wp2 == w1) return 0;
    }
    for (int e = 0; e < d->nexclude; e++) {
        int e1 = d->exclude_body[2 * e], e2 = d->exclude_body[2 * e + 1];
        if ((e1 == b1 && e2 == b2) || (e1 == b2 && e2 == b1)) return 0;
    }
    if (!((s->geom_contype[g1] & s->geom_conaffinity[g2]) || (s->geom_contype[g2] & s->geom_conaffinity[g1]))) return 0;
    if (s->geom_type[g1] == MOPA_GEOM_PLANE && s->geom_type[g2] == MOPA_GEOM_PLANE) return 0;
    return 1;
}

void *orc_scene_create(const mopa_model_desc *d, const int32_t *ignored_pairs, int nignored, double contact_threshold) {
    orc_scene *s = (orc_scene *)calloc(1, sizeof(orc_scene));
    s->nq = d->nq; s->nbody = d->nbody; s->njnt = d->njnt; s->ngeom = d->ngeom; s->nsite = d->nsite;
    s->body_parentid = dupi(d->body_parentid, d->nbody);
    s->body_weldid = dupi(d->body_weldid, d->nbody);
    s->body_jntadr = dupi(d->body_jntadr, d->nbody);
    s->body_jntnum = dupi(d->body_jntnum, d->nbody);
    s->body_pos = dupr(d->body_pos, 3 * d->nbody);
    s->body_quat = dupr(d->body_quat, 4 * d->nbody);
    s->jnt_type = dupi(d->jnt_type, d->njnt);
    s->jnt_qposadr = dupi(d->jnt_qposadr, d->njnt);
    s->jnt_bodyid = dupi(d->jnt_bodyid, d->njnt);
    s->jnt_limited = dupi(d->jnt_limited, d->njnt);
    s->jnt_pos = dupr(d->jnt_pos, 3 * d->njnt);
    s->jnt_axis = dupr(d->jnt_axis, 3 * d->njnt);
    s->jnt_range = dupr(d->jnt_range, 2 * d->njnt);
    s->qpos0 = dupr(d->qpos0, d->nq);
    s->geom_type = dupi(d->geom_type, d->ngeom);
    s->geom_bodyid = dupi(d->geom_bodyid, d->ngeom);
    s->geom_contype = dupi(d->geom_contype, d->ngeom);
    s->geom_conaffinity = dupi(d->geom_conaffinity, d->ngeom);
    s->geom_pos = dupr(d->geom_pos, 3 * d->ngeom);
    s->geom_size = dupr(d->geom_size, 3 * d->ngeom);
    s->geom_margin = dupr(d->geom_margin, d->ngeom);
    s->geom_rbound = dupr(d->geom_rbound, d->ngeom);
    s->geom_dataid = dupi(d->geom_dataid, d->ngeom);
    s->mesh_vertadr = dupi(d->mesh_vertadr, d->nmesh);
    s->mesh_vertnum = dupi(d->mesh_vertnum, d->nmesh);
    s->mesh_vert = dupr(d->mesh_vert, 3 * d->nmeshvert);
    s->geom_mat = (R *)malloc(sizeof(R) * 9 * (d->ngeom + 1));
    for (int g = 0; g < d->ngeom; g++) {
        R q[4] = {(R)d->geom_quat[4 * g], (R)d->geom_quat[4 * g + 1], (R)d->geom_quat[4 * g + 2], (R)d->geom_quat[4 * g + 3]};
        q2m(s->geom_mat + 9 * g, q);
    }
    s->site_bodyid = dupi(d->site_bodyid, d->nsite);
    s->site_pos = dupr(d->site_pos, 3 * d->nsite);
    s->site_mat = (R *)malloc(sizeof(R) * 9 * (d->nsite + 1));
    for (int g = 0; g < d->nsite; g++) {
        R q[4] = {(R)d->site_quat[4 * g], (R)d->site_quat[4 * g + 1], (R)d->site_quat[4 * g + 2], (R)d->site_quat[4 * g + 3]};
        q2m(s->site_mat + 9 * g, q);
    }
    s->threshold = (R)contact_threshold;
    /* canonical pair list: g1 < g2 lexicographic, filters of B.3, minus ignored_contacts
       (mujoco_ompl_interface.cpp:955-959 skips them whatever their distance) */
    s->pair_g1 = (int *)malloc(sizeof(int) * d->ngeom * d->ngeom);
    s->pair_g2 = (int *)malloc(sizeof(int) * d->ngeom * d->ngeom);
    for (int g1 = 0; g1 < d->ngeom; g1++)
        for (int g2 = g1 + 1; g2 < d->ngeom; g2++) {
            if (!pair_allowed(s, d, g1, g2)) continue;
            int ign = 0;
            for (int k = 0; k < nignored; k++) {
                int a = ignored_pairs[2 * k], b = ignored_pairs[2 * k + 1];
                /* the reference compares against make_ordered_pair(geom1, geom2) = (min,max) */
                if (a == g1 && b == g2) ign = 1;
            }
            if (ign) continue;
            s->pair_g1[s->npair] = g1;
            s->pair_g2[s->npair] = g2;
            s->npair++;
        }
    s->xpos = (R *)malloc(sizeof(R) * 3 * d->nbody);
    s->xquat = (R *)malloc(sizeof(R) * 4 * d->nbody);
    s->xmat = (R *)malloc(sizeof(R) * 9 * d->nbody);
    s->gpos = (R *)malloc(sizeof(R) * 3 * (d->ngeom + 1));
    s->gmat = (R *)malloc(sizeof(R) * 9 * (d->ngeom + 1));
    return s;
}

void orc_scene_destroy(void *h) {
    orc_scene *s = (orc_scene *)h;
    if (!s) return;
    free(s->body_parentid); free(s->body_weldid); free(s->body_jntadr); free(s->body_jntnum);
    free(s->body_pos); free(s->body_quat); free(s->jnt_type); free(s->jnt_qposadr); free(s->jnt_bodyid);
    free(s->jnt_limited); free(s->jnt_pos); free(s->jnt_axis); free(s->jnt_range); free(s->qpos0);
    free(s->geom_type); free(s->geom_bodyid); free(s->geom_contype); free(s->geom_conaffinity);
    free(s->geom_pos); free(s->geom_mat); free(s->geom_size); free(s->geom_margin); free(s->geom_rbound);
    free(s->site_bodyid); free(s->site_pos); free(s->site_mat);
    free(s->geom_dataid); free(s->mesh_vertadr); free(s->mesh_vertnum); free(s->mesh_vert);
    free(s->pair_g1); free(s->pair_g2); free(s->xpos); free(s->xquat); free(s->xmat); free(s->gpos); free(s->gmat);
    free(s);
}

int orc_scene_npair(void *h) { return ((orc_scene *)h)->npair; }
void orc_scene_pairs(void *h, int32_t *g1, int32_t *g2) {
    orc_scene *s = (orc_scene *)h;
    for (int i = 0; i < s->npair; i++) { g1[i] = s->pair_g1[i]; g2[i] = s->pair_g2[i]; }
}

/* ---------------------------------------------------------------- forward kinematics (B.1) */
static void fk(orc_scene *s, const R *qpos) {
    R *xpos = s->xpos, *xquat = s->xquat, *xmat = s->xmat;
    xpos[0] = xpos[1] = xpos[2] = 0;
    xquat[0] = 1; xquat[1] = xquat[2] = xquat[3] = 0;
    q2m(xmat, xquat);
    for (int b = 1; b < s->nbody; b++) {
        int p = s->body_parentid[b];
        R pos[3], quat[4], M[9], t[3];
        mulMV(t, xmat + 9 * p, s->body_pos + 3 * b);
        add3(pos, xpos + 3 * p, t);
        qmul(quat, xquat + 4 * p, s->body_quat + 4 * b);
        for (int k = 0; k < s->body_jntnum[b]; k++) {
            int j = s->body_jntadr[b] + k;
            int a = s->jnt_qposadr[j];
            if (s->jnt_type[j] == MOPA_JNT_FREE) {
                pos[0] = qpos[a]; pos[1] = qpos[a + 1]; pos[2] = qpos[a + 2];
                R w = qpos[a + 3], x = qpos[a + 4], y = qpos[a + 5], z = qpos[a + 6];
                R n = RSQRT_(MAD(z, z, MAD(y, y, MAD(x, x, w * w))));
                quat[0] = w / n; quat[1] = x / n; quat[2] = y / n; quat[3] = z / n;
            } else if (s->jnt_type[j] == MOPA_JNT_SLIDE) {
                R ax[3];
                q2m(M, quat);
                mulMV(ax, M, s->jnt_axis + 3 * j);
                R dq = qpos[a] - s->qpos0[a];
                pos[0] = MAD(ax[0], dq, pos[0]); pos[1] = MAD(ax[1], dq, pos[1]); pos[2] = MAD(ax[2], dq, pos[2]);
            } else if (s->jnt_type[j] == MOPA_JNT_HINGE) {
                R anchor[3], sn, cs, ql[4], qn[4];
                q2m(M, quat);
                mulMV(t, M, s->jnt_pos + 3 * j);
                add3(anchor, pos, t);
                sincos_r((qpos[a] - s->qpos0[a]) * RC(0.5), &sn, &cs);
                ql[0] = cs; ql[1] = sn * s->jnt_axis[3 * j]; ql[2] = sn * s->jnt_axis[3 * j + 1]; ql[3] = sn * s->jnt_axis[3 * j + 2];
                qmul(qn, quat, ql);
                memcpy(quat, qn, sizeof(qn));
                q2m(M, quat);
                mulMV(t, M, s->jnt_pos + 3 * j);
                sub3(pos, anchor, t);
            }
        }
        memcpy(xpos + 3 * b, pos, sizeof(pos));
        memcpy(xquat + 4 * b, quat, sizeof(quat));
        q2m(xmat + 9 * b, quat);
    }
    for (int g = 0; g < s->ngeom; g++) {
        int b = s->geom_bodyid[g];
        R t[3];
        mulMV(t, xmat + 9 * b, s->geom_pos + 3 * g);
        add3(s->gpos + 3 * g, xpos + 3 * b, t);
        mulMM(s->gmat + 9 * g, xmat + 9 * b, s->geom_mat + 9 * g);
    }
}

/* ---------------------------------------------------------------- analytic narrowphase (B.5) */
/* geom frame axis k (column k of the row-major rotation) */
static inline void col3(R *a, const R *M, int k) { a[0] = M[k]; a[1] = M[3 + k]; a[2] = M[6 + k]; }

static R plane_sphere(const R *pp, const R *pm, const R *c, R r) {
    R n[3], d[3];
    col3(n, pm, 2); sub3(d, c, pp);
    return dot3(n, d) - r;
}
static R plane_capsule(const R *pp, const R *pm, const R *c, const R *m, const R *sz) {
    R n[3], a[3], d[3];
    col3(n, pm, 2); col3(a, m, 2); sub3(d, c, pp);
    R hc = dot3(n, d), ha = dot3(n, a) * sz[1];
    return (hc - FABS_(ha)) - sz[0];
}
static R plane_cylinder(const R *pp, const R *pm, const R *c, const R *m, const R *sz) {
    R n[3], a[3], d[3];
    col3(n, pm, 2); col3(a, m, 2); sub3(d, c, pp);
    R hc = dot3(n, d), na = dot3(n, a);
    R s2 = FMAX_(RC(0.0), MAD(-na, na, RC(1.0)));
    return (hc - FABS_(na) * sz[1]) - sz[0] * RSQRT_(s2);
}
static R plane_box(const R *pp, const R *pm, const R *c, const R *m, const R *sz) {
    R n[3], d[3], l[3];
    col3(n, pm, 2); sub3(d, c, pp);
    mulMTV(l, m, n); /* plane normal in box frame */
    R ext = MAD(FABS_(l[2]), sz[2], MAD(FABS_(l[1]), sz[1], FABS_(l[0]) * sz[0]));
    return dot3(n, d) - ext;
}
/* index of the hull vertex that is extreme along the LOCAL direction l (first maximum) */
static int hull_extreme(const R *vert, int n, const R *l) {
    int best = 0;
    R bd = dot3(vert, l);
    for (int i = 1; i < n; i++) {
        R di = dot3(vert + 3 * i, l);
        if (di > bd) { bd = di; best = i; }
    }
    return best;
}

/* plane - convex hull: lowest hull vertex along the plane normal */
static R plane_mesh(const R *pp, const R *pm, const R *c, const R *m, const R *vert, int nvert) {
    R n[3], d[3], l[3], nl[3];
    col3(n, pm, 2);
    sub3(d, c, pp);
    mulMTV(l, m, n);
    nl[0] = -l[0]; nl[1] = -l[1]; nl[2] = -l[2];
    return dot3(n, d) + dot3(vert + 3 * hull_extreme(vert, nvert, nl), l);
}
static R sphere_sphere(const R *c1, R r1, const R *c2, R r2) {
    R d[3];
    sub3(d, c2, c1);
    return (len3(d) - r1) - r2;
}
/* squared distance point p -> segment [c - a*h, c + a*h] */
static R point_seg(const R *p, const R *c, const R *a, R h) {
    R d[3], w[3];
    sub3(d, p, c);
    R t = dot3(d, a);
    t = FMIN_(h, FMAX_(-h, t));
    w[0] = MAD(-t, a[0], d[0]); w[1] = MAD(-t, a[1], d[1]); w[2] = MAD(-t, a[2], d[2]);
    return len3(w);
}
static R sphere_capsule(const R *c1, R r1, const R *c2, const R *m2, const R *sz2) {
    R a[3];
    col3(a, m2, 2);
    return (point_seg(c1, c2, a, sz2[1]) - r1) - sz2[0];
}
static R capsule_capsule(const R *c1, const R *m1, const R *sz1, const R *c2, const R *m2, const R *sz2) {
    /* closest points of segments P(s)=c1+s*a1 (|s|<=h1), Q(t)=c2+t*a2 (|t|<=h2); unit axes */
    R a1[3], a2[3], r[3], w[3];
    col3(a1, m1, 2); col3(a2, m2, 2);
    R h1 = sz1[1], h2 = sz2[1];
    sub3(r, c1, c2);
    R b = dot3(a1, a2), c = dot3(a1, r), f = dot3(a2, r);
    R den = MAD(-b, b, RC(1.0));
    R s, t;
    if (den > RC(1e-6)) {
        s = MAD(b, f, -c) / den;
        s = FMIN_(h1, FMAX_(-h1, s));
    } else
        s = RC(0.0);
    t = MAD(b, s, f);
    if (t < -h2) { t = -h2; s = FMIN_(h1, FMAX_(-h1, MAD(b, t, -c))); }
    else if (t > h2) { t = h2; s = FMIN_(h1, FMAX_(-h1, MAD(b, t, -c))); }
    w[0] = MAD(-t, a2[0], MAD(s, a1[0], r[0]));
    w[1] = MAD(-t, a2[1], MAD(s, a1[1], r[1]));
    w[2] = MAD(-t, a2[2], MAD(s, a1[2], r[2]));
    return (len3(w) - sz1[0]) - sz2[0];
}
static R sphere_cylinder(const R *c1, R r1, const R *c2, const R *m2, const R *sz2) {
    R a[3], d[3], w[3];
    col3(a, m2, 2); sub3(d, c1, c2);
    R z = dot3(d, a);
    w[0] = MAD(-z, a[0], d[0]); w[1] = MAD(-z, a[1], d[1]); w[2] = MAD(-z, a[2], d[2]);
    R dr = len3(w) - sz2[0], dz = FABS_(z) - sz2[1];
    R core;
    if (dr <= 0 && dz <= 0) core = FMAX_(dr, dz);
    else if (dz <= 0) core = dr;
    else if (dr <= 0) core = dz;
    else core = RSQRT_(MAD(dr, dr, dz * dz));
    return core - r1;
}
static R sphere_box(const R *c1, R r1, const R *c2, const R *m2, const R *sz2) {
    R d[3], p[3], e[3];
    sub3(d, c1, c2);
    mulMTV(p, m2, d);
    e[0] = FABS_(p[0]) - sz2[0]; e[1] = FABS_(p[1]) - sz2[1]; e[2] = FABS_(p[2]) - sz2[2];
    R core;
    if (e[0] <= 0 && e[1] <= 0 && e[2] <= 0) core = FMAX_(e[0], FMAX_(e[1], e[2]));
    else {
        R o[3] = {FMAX_(e[0], RC(0.0)), FMAX_(e[1], RC(0.0)), FMAX_(e[2], RC(0.0))};
        core = len3(o);
    }
    return core - r1;
}
/* 15-axis separating-axis test.  Returns max over axes of the signed separation:
   > 0 disjoint (lower bound of the distance), < 0 minus the penetration depth. */
static R box_box(const R *c1, const R *m1, const R *sz1, const R *c2, const R *m2, const R *sz2) {
    R d[3], T[3], Rm[9], A[9];
    sub3(d, c2, c1);
    mulMTV(T, m1, d);
    /* Rm = m1^T m2 */
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            Rm[3 * i + j] = MAD(m1[6 + i], m2[6 + j], MAD(m1[3 + i], m2[3 + j], m1[i] * m2[j]));
            A[3 * i + j] = FABS_(Rm[3 * i + j]);
        }
    R best = -ORC_BIG;
    for (int i = 0; i < 3; i++) { /* faces of box 1 */
        R rb = MAD(A[3 * i + 2], sz2[2], MAD(A[3 * i + 1], sz2[1], A[3 * i] * sz2[0]));
        R sep = (FABS_(T[i]) - sz1[i]) - rb;
        best = FMAX_(best, sep);
    }
    for (int j = 0; j < 3; j++) { /* faces of box 2 */
        R ra = MAD(A[6 + j], sz1[2], MAD(A[3 + j], sz1[1], A[j] * sz1[0]));
        R tp = MAD(T[2], Rm[6 + j], MAD(T[1], Rm[3 + j], T[0] * Rm[j]));
        R sep = (FABS_(tp) - ra) - sz2[j];
        best = FMAX_(best, sep);
    }
    for (int i = 0; i < 3; i++) { /* edge x edge */
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; j++) {
            int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            /* axis L = e1_i x e2_j expressed in frame 1: |L|^2 = 1 - Rm[i][j]^2 */
            R l2 = MAD(-Rm[3 * i + j], Rm[3 * i + j], RC(1.0));
            if (l2 < RC(1e-6)) continue; /* parallel edges: covered by the face axes */
            R ra = MAD(sz1[i2], A[3 * i1 + j], sz1[i1] * A[3 * i2 + j]);
            R rb = MAD(sz2[j2], A[3 * i + j1], sz2[j1] * A[3 * i + j2]);
            R tp = MAD(T[i2], Rm[3 * i1 + j], -(T[i1] * Rm[3 * i2 + j]));
            R sep = ((FABS_(tp) - ra) - rb) / RSQRT_(l2);
            best = FMAX_(best, sep);
        }
    }
    return best;
}

/* ---------------------------------------------------------------- generic convex: MPR */
typedef struct { int type; const R *pos, *mat, *size; const R *vert; int nvert; } cvx;

/* support point of a convex geom in world direction dir (unit length).
   Capsules and cylinders only use their axis (z column of the frame). */
static void support(R *out, const cvx *g, const R *dir) {
    if (g->type == MOPA_GEOM_MESH) { /* convex hull of the mesh: extreme vertex (libccd support of mjc_Convex) */
        R l[3], w[3];
        mulMTV(l, g->mat, dir);
        mulMV(w, g->mat, g->vert + 3 * hull_extreme(g->vert, g->nvert, l));
        add3(out, g->pos, w);
        return;
    }
    if (g->type == MOPA_GEOM_SPHERE) {
        out[0] = MAD(dir[0], g->size[0], g->pos[0]);
        out[1] = MAD(dir[1], g->size[0], g->pos[1]);
        out[2] = MAD(dir[2], g->size[0], g->pos[2]);
        return;
    }
    if (g->type == MOPA_GEOM_BOX) {
        R l[3], p[3], w[3];
        mulMTV(l, g->mat, dir);
        p[0] = l[0] >= 0 ? g->size[0] : -g->size[0];
        p[1] = l[1] >= 0 ? g->size[1] : -g->size[1];
        p[2] = l[2] >= 0 ? g->size[2] : -g->size[2];
        mulMV(w, g->mat, p);
        add3(out, g->pos, w);
        return;
    }
    R a[3];
    col3(a, g->mat, 2);
    R z = dot3(dir, a);
    R hs = z >= 0 ? g->size[1] : -g->size[1];
    if (g->type == MOPA_GEOM_CYLINDER) {
        R w[3] = {MAD(-z, a[0], dir[0]), MAD(-z, a[1], dir[1]), MAD(-z, a[2], dir[2])};
        R n = len3(w);
        R k = n > RC(1e-12) ? g->size[0] / n : RC(0.0);
        out[0] = MAD(w[0], k, MAD(a[0], hs, g->pos[0]));
        out[1] = MAD(w[1], k, MAD(a[1], hs, g->pos[1]));
        out[2] = MAD(w[2], k, MAD(a[2], hs, g->pos[2]));
    } else { /* capsule */
        out[0] = MAD(dir[0], g->size[0], MAD(a[0], hs, g->pos[0]));
        out[1] = MAD(dir[1], g->size[0], MAD(a[1], hs, g->pos[1]));
        out[2] = MAD(dir[2], g->size[0], MAD(a[2], hs, g->pos[2]));
    }
}
/* support of the Minkowski difference g1 - g2 */
static void msupport(R *v, const cvx *g1, const cvx *g2, const R *dir) {
    R a[3], b[3], nd[3] = {-dir[0], -dir[1], -dir[2]};
    support(a, g1, dir);
    support(b, g2, nd);
    sub3(v, a, b);
}
static inline int normalize3(R *v) {
    R n = len3(v);
    if (n < RC(1e-30)) return 0;
    v[0] /= n; v[1] /= n; v[2] /= n;
    return 1;
}
#define MPR_EPS RC(1.1920929e-07)
/* tolerance / iteration cap; adjustable for convergence studies (orc_set_mpr) */
static R MPR_TOL = RC(1e-6);
static int MPR_MAXIT = 50;
static long g_mpr_calls = 0, g_mpr_iters = 0, g_mpr_hits = 0;
void orc_set_mpr(int maxit, double tol) { MPR_MAXIT = maxit; MPR_TOL = (R)tol; }
void orc_mpr_stats(long *out, int reset) {
    out[0] = g_mpr_calls; out[1] = g_mpr_iters; out[2] = g_mpr_hits;
    if (reset) g_mpr_calls = g_mpr_iters = g_mpr_hits = 0;
}
static inline int is_zero(R x) { return FABS_(x) < MPR_EPS; }

/* squared distance from the origin to triangle (a,b,c) (region classification) */
static R origin_tri_dist2(const R *a, const R *b, const R *c) {
    R ab[3], ac[3], ap[3] = {-a[0], -a[1], -a[2]};
    sub3(ab, b, a); sub3(ac, c, a);
    R d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    if (d1 <= 0 && d2 <= 0) return dot3(a, a);
    R bp[3] = {-b[0], -b[1], -b[2]};
    R d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    if (d3 >= 0 && d4 <= d3) return dot3(b, b);
    R vc = MAD(d1, d4, -(d3 * d2));
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        R v = d1 / (d1 - d3);
        R q[3] = {MAD(v, ab[0], a[0]), MAD(v, ab[1], a[1]), MAD(v, ab[2], a[2])};
        return dot3(q, q);
    }
    R cp[3] = {-c[0], -c[1], -c[2]};
    R d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) return dot3(c, c);
    R vb = MAD(d5, d2, -(d1 * d6));
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        R w = d2 / (d2 - d6);
        R q[3] = {MAD(w, ac[0], a[0]), MAD(w, ac[1], a[1]), MAD(w, ac[2], a[2])};
        return dot3(q, q);
    }
    R va = MAD(d3, d6, -(d5 * d4));
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        R w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        R bc[3];
        sub3(bc, c, b);
        R q[3] = {MAD(w, bc[0], b[0]), MAD(w, bc[1], b[1]), MAD(w, bc[2], b[2])};
        return dot3(q, q);
    }
    /* interior: distance to the plane */
    R n[3];
    cross3(n, ab, ac);
    R nn = dot3(n, n);
    if (nn < RC(1e-30)) return dot3(a, a);
    R k = dot3(n, a);
    return (k * k) / nn;
}

/* Returns 1 and *depth when the shapes intersect, 0 otherwise. */
static int mpr_penetration(const cvx *g1, const cvx *g2, R *depth) {
    R v0[3], v1[3], v2[3], v3[3], v4[3], dir[3], va[3], vb[3];
    R dot;
    g_mpr_calls++;
    /* phase 1: portal discovery */
    sub3(v0, g1->pos, g2->pos);
    if (v0[0] == 0 && v0[1] == 0 && v0[2] == 0) v0[0] = MPR_EPS * RC(10.0);
    dir[0] = -v0[0]; dir[1] = -v0[1]; dir[2] = -v0[2];
    normalize3(dir);
    msupport(v1, g1, g2, dir);
    dot = dot3(v1, dir);
    if (is_zero(dot) || dot < 0) return 0;
    cross3(dir, v0, v1);
    if (is_zero(dot3(dir, dir))) {
        if (v1[0] == 0 && v1[1] == 0 && v1[2] == 0) { *depth = 0; return 1; } /* touching at v1 */
        *depth = len3(v1); /* origin on the segment v0-v1 */
        return 1;
    }
    normalize3(dir);
    msupport(v2, g1, g2, dir);
    dot = dot3(v2, dir);
    if (is_zero(dot) || dot < 0) return 0;
    sub3(va, v1, v0); sub3(vb, v2, v0);
    cross3(dir, va, vb);
    normalize3(dir);
    dot = dot3(dir, v0);
    if (dot > 0) {
        R t[3];
        cpy3(t, v1); cpy3(v1, v2); cpy3(v2, t);
        dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2];
    }
    int it = 0;
    for (;;) {
        if (++it > MPR_MAXIT) return 0;
        msupport(v3, g1, g2, dir);
        dot = dot3(v3, dir);
        if (is_zero(dot) || dot < 0) return 0;
        int cont = 0;
        cross3(va, v1, v3);
        dot = dot3(va, v0);
        if (dot < 0 && !is_zero(dot)) { cpy3(v2, v3); cont = 1; }
        if (!cont) {
            cross3(va, v3, v2);
            dot = dot3(va, v0);
            if (dot < 0 && !is_zero(dot)) { cpy3(v1, v3); cont = 1; }
        }
        if (!cont) break;
        sub3(va, v1, v0); sub3(vb, v2, v0);
        cross3(dir, va, vb);
        normalize3(dir);
    }
    /* phase 2: portal refinement; phase 3: penetration depth */
    int inside = 0;
    for (it = 0;; it++) {
        sub3(va, v2, v1); sub3(vb, v3, v1);
        cross3(dir, va, vb);
        normalize3(dir);
        if (!inside) {
            dot = dot3(dir, v1);
            if (is_zero(dot) || dot > 0) inside = 1; /* portal encapsulates the origin */
        }
        msupport(v4, g1, g2, dir);
        R dv4 = dot3(v4, dir);
        R dmin = FMIN_(dv4 - dot3(v1, dir), FMIN_(dv4 - dot3(v2, dir), dv4 - dot3(v3, dir)));
        int reached = (dmin <= MPR_TOL);
        if (!inside) {
            if (!(is_zero(dv4) || dv4 > 0) || reached || it >= MPR_MAXIT) return 0;
        } else if (reached || it >= MPR_MAXIT) {
            *depth = RSQRT_(origin_tri_dist2(v1, v2, v3));
            g_mpr_iters += it; g_mpr_hits++;
            return 1;
        }
        /* expand the portal with v4 */
        cross3(va, v4, v0);
        dot = dot3(v1, va);
        if (dot > 0) {
            dot = dot3(v2, va);
            if (dot > 0) cpy3(v1, v4); else cpy3(v3, v4);
        } else {
            dot = dot3(v3, va);
            if (dot > 0) cpy3(v2, v4); else cpy3(v1, v4);
        }
    }
}

/* signed distance of one candidate pair; ORC_BIG when the pair cannot be in contact */
static R pair_dist(const orc_scene *s, int g1, int g2) {
    int t1 = s->geom_type[g1], t2 = s->geom_type[g2];
    if (t1 > t2) { int t = g1; g1 = g2; g2 = t; t = t1; t1 = t2; t2 = t; }
    const R *c1 = s->gpos + 3 * g1, *c2 = s->gpos + 3 * g2;
    const R *m1 = s->gmat + 9 * g1, *m2 = s->gmat + 9 * g2;
    const R *z1 = s->geom_size + 3 * g1, *z2 = s->geom_size + 3 * g2;
    R margin = FMAX_(s->geom_margin[g1], s->geom_margin[g2]);
    if (t1 != MOPA_GEOM_PLANE) { /* bounding-sphere filter (planes are unbounded) */
        R d[3];
        sub3(d, c2, c1);
        R bound = (s->geom_rbound[g1] + s->geom_rbound[g2]) + margin;
        if (dot3(d, d) > bound * bound) return ORC_BIG;
    }
    switch (t1) {
    case MOPA_GEOM_PLANE:
        if (t2 == MOPA_GEOM_SPHERE) return plane_sphere(c1, m1, c2, z2[0]);
        if (t2 == MOPA_GEOM_CAPSULE) return plane_capsule(c1, m1, c2, m2, z2);
        if (t2 == MOPA_GEOM_CYLINDER) return plane_cylinder(c1, m1, c2, m2, z2);
        if (t2 == MOPA_GEOM_BOX) return plane_box(c1, m1, c2, m2, z2);
        if (t2 == MOPA_GEOM_MESH) {
            int me = s->geom_dataid[g2];
            return plane_mesh(c1, m1, c2, m2, s->mesh_vert + 3 * s->mesh_vertadr[me], s->mesh_vertnum[me]);
        }
        return ORC_BIG;
    case MOPA_GEOM_SPHERE:
        if (t2 == MOPA_GEOM_SPHERE) return sphere_sphere(c1, z1[0], c2, z2[0]);
        if (t2 == MOPA_GEOM_CAPSULE) return sphere_capsule(c1, z1[0], c2, m2, z2);
        if (t2 == MOPA_GEOM_CYLINDER) return sphere_cylinder(c1, z1[0], c2, m2, z2);
        if (t2 == MOPA_GEOM_BOX) return sphere_box(c1, z1[0], c2, m2, z2);
        if (t2 != MOPA_GEOM_MESH) return ORC_BIG;
        break;
    case MOPA_GEOM_CAPSULE:
        if (t2 == MOPA_GEOM_CAPSULE) return capsule_capsule(c1, m1, z1, c2, m2, z2);
        break;
    case MOPA_GEOM_BOX:
        if (t2 == MOPA_GEOM_BOX) return box_box(c1, m1, z1, c2, m2, z2);
        break;
    default:
        break;
    }
    /* everything else, including every pair with a mesh geom (convex hull), goes through MPR */
    cvx a = {t1, c1, m1, z1, 0, 0}, b = {t2, c2, m2, z2, 0, 0};
    if (t1 == MOPA_GEOM_MESH) { int me = s->geom_dataid[g1]; a.vert = s->mesh_vert + 3 * s->mesh_vertadr[me]; a.nvert = s->mesh_vertnum[me]; }
    if (t2 == MOPA_GEOM_MESH) { int me = s->geom_dataid[g2]; b.vert = s->mesh_vert + 3 * s->mesh_vertadr[me]; b.nvert = s->mesh_vertnum[me]; }
    R depth;
    if (mpr_penetration(&a, &b, &depth)) return -depth;
    return ORC_BIG;
}

/* result word: bit0 = valid; bits 8.. = 1 + index (canonical pair list) of the first
   offending pair, 0 when valid. */
int orc_is_valid(void *h, const double *qpos, int n, uint32_t *result, double *min_dist) {
    orc_scene *s = (orc_scene *)h;
    R *q = (R *)malloc(sizeof(R) * s->nq);
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < s->nq; k++) q[k] = (R)qpos[(size_t)i * s->nq + k];
        fk(s, q);
        uint32_t first = 0;
        R md = ORC_BIG;
        for (int p = 0; p < s->npair; p++) {
            R d = pair_dist(s, s->pair_g1[p], s->pair_g2[p]);
            if (d < md) md = d;
            if (d <= s->threshold && !first) first = (uint32_t)p + 1;
        }
        result[i] = first ? (first << 8) : 1u;
        if (min_dist) min_dist[i] = (double)md;
    }
    free(q);
    return 0;
}

/* per-pair signed distances for one state (debug / unit tests) */
int orc_pair_dists(void *h, const double *qpos, double *dist) {
    orc_scene *s = (orc_scene *)h;
    R *q = (R *)malloc(sizeof(R) * s->nq);
    for (int k = 0; k < s->nq; k++) q[k] = (R)qpos[k];
    fk(s, q);
    for (int p = 0; p < s->npair; p++) dist[p] = (double)pair_dist(s, s->pair_g1[p], s->pair_g2[p]);
    free(q);
    return 0;
}

/* world frames of every body / geom / site for one state */
int orc_fk(void *h, const double *qpos, double *body_xpos, double *body_xmat, double *geom_xpos, double *geom_xmat,
           double *site_xpos, double *site_xmat) {
    orc_scene *s = (orc_scene *)h;
    R *q = (R *)malloc(sizeof(R) * s->nq);
    for (int k = 0; k < s->nq; k++) q[k] = (R)qpos[k];
    fk(s, q);
    for (int i = 0; i < 3 * s->nbody; i++) body_xpos[i] = s->xpos[i];
    for (int i = 0; i < 9 * s->nbody; i++) body_xmat[i] = s->xmat[i];
    for (int i = 0; i < 3 * s->ngeom; i++) geom_xpos[i] = s->gpos[i];
    for (int i = 0; i < 9 * s->ngeom; i++) geom_xmat[i] = s->gmat[i];
    for (int g = 0; g < s->nsite; g++) {
        int b = s->site_bodyid[g];
        R t[3], p[3], M[9];
        mulMV(t, s->xmat + 9 * b, s->site_pos + 3 * g);
        add3(p, s->xpos + 3 * b, t);
        mulMM(M, s->xmat + 9 * b, s->site_mat + 9 * g);
        if (site_xpos) for (int k = 0; k < 3; k++) site_xpos[3 * g + k] = p[k];
        if (site_xmat) for (int k = 0; k < 9; k++) site_xmat[9 * g + k] = M[k];
    }
    free(q);
    return 0;
}

/* pairwise primitive test entry for unit tests: frames given explicitly */
double orc_primitive_dist(int t1, const double *p1, const double *m1, const double *s1, int t2, const double *p2,
                          const double *m2, const double *s2) {
    orc_scene s;
    memset(&s, 0, sizeof(s));
    int gt[2] = {t1, t2};
    R gp[6], gm[18], gs[6], mg[2] = {0, 0}, rb[2] = {ORC_BIG * RC(1e-6), ORC_BIG * RC(1e-6)};
    for (int k = 0; k < 3; k++) { gp[k] = (R)p1[k]; gp[3 + k] = (R)p2[k]; gs[k] = (R)s1[k]; gs[3 + k] = (R)s2[k]; }
    for (int k = 0; k < 9; k++) { gm[k] = (R)m1[k]; gm[9 + k] = (R)m2[k]; }
    s.geom_type = gt; s.gpos = gp; s.gmat = gm; s.geom_size = gs; s.geom_margin = mg; s.geom_rbound = rb;
    return (double)pair_dist(&s, 0, 1);
}

/* single-state predicate for the planner oracle (orc_plan.c); q has nq entries */
int orc_valid_one(void *h, const R *q) {
    orc_scene *s = (orc_scene *)h;
    fk(s, q);
    for (int p = 0; p < s->npair; p++)
        if (pair_dist(s, s->pair_g1[p], s->pair_g2[p]) <= s->threshold) return 0;
    return 1;
}
int orc_scene_nq(void *h) { return ((orc_scene *)h)->nq; }
