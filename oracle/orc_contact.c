/* TEST INFRASTRUCTURE — CPU oracle for the MoPA-RL hot path.  Not part of the product.
 *
 * Contact generation for the physics-step oracle: candidate geom pairs (filters of SURVEY.md
 * App. B.3, evaluated by the host in dynmodel.py) -> bounding-sphere cull with margin ->
 * contact points (position, normal from geom 1 to geom 2, signed distance) -> constraint rows
 * (normal + two tangents, Jacobians from the joint motion axes).
 *
 * MuJoCo 2.0's collision functions are closed source (PARITY UNPINNED).  Restated here:
 *   plane-{sphere,capsule,box}, sphere-sphere, sphere-capsule, capsule-capsule, sphere-box,
 *   sphere-cylinder: closed form;
 *   box-box: 15-axis SAT, then face clipping (<= 4 points) or the edge-edge closest points;
 *   capsule-{box,cylinder}, cylinder-{box,cylinder}: closest points of the convex cores by
 *   alternating projections (fixed 16 sweeps); a cylinder is treated as its core shrunk by
 *   rho = min(5 mm, half of its smaller dimension) swept by a sphere of radius rho (rims rounded
 *   by at most (sqrt(2)-1) rho), a box facing a cylinder likewise.  MuJoCo uses libccd MPR here.
 * Contact parameters combine per pair as MuJoCo does: margin = max, friction = max, solref /
 * solimp = mean (equal solmix).  Only sliding friction is modelled (condim 3).
 */
#include <math.h>
#include <string.h>

#include "orc_dyn.h"

typedef struct { double pos[3], n[3], dist; } cpoint;
typedef struct { const double *c, *R, *size; int type; } cgeom; /* world centre, rotation (row-major), size */

static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double *r, const double *a, const double *b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static void colk(double *a, const double *R, int k) { a[0] = R[k]; a[1] = R[3 + k]; a[2] = R[6 + k]; }
static void to_local(double *l, const cgeom *g, const double *p) {
    double d[3] = {p[0] - g->c[0], p[1] - g->c[1], p[2] - g->c[2]};
    for (int k = 0; k < 3; k++) l[k] = g->R[k] * d[0] + g->R[3 + k] * d[1] + g->R[6 + k] * d[2];
}
static void to_world(double *p, const cgeom *g, const double *l) {
    for (int k = 0; k < 3; k++) p[k] = g->c[k] + g->R[3 * k] * l[0] + g->R[3 * k + 1] * l[1] + g->R[3 * k + 2] * l[2];
}
static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* contact of a sphere (centre s, radius r) against point q on / in the other shape:
   `inside` != 0 means s is inside the other shape and `nout` is the outward normal at q */
static int sphere_vs_point(cpoint *cp, const double *s, double r, const double *q, int inside, const double *nout, double margin,
                           int sphere_is_g1) {
    double d[3] = {s[0] - q[0], s[1] - q[1], s[2] - q[2]}, len = sqrt(dot3(d, d)), n[3], dist;
    if (inside) { memcpy(n, nout, sizeof(n)); dist = -len - r; }
    else {
        if (len < 1e-12) return 0;
        for (int k = 0; k < 3; k++) n[k] = d[k] / len; /* from the other shape towards the sphere */
        dist = len - r;
    }
    if (dist >= margin) return 0;
    cp->dist = dist;
    for (int k = 0; k < 3; k++) {
        cp->n[k] = sphere_is_g1 ? -n[k] : n[k]; /* normal from geom 1 to geom 2 */
        cp->pos[k] = q[k] + n[k] * 0.5 * dist;
    }
    return 1;
}

/* closest point of a box (half sizes h, shrunk by rho) to local point l; returns 1 if l is inside */
static int box_closest(double *q, const double *l, const double *h, double rho, double *nout_local) {
    int inside = 1, best = 0;
    double bestd = 1e30;
    for (int k = 0; k < 3; k++) {
        double hk = h[k] - rho;
        q[k] = clampd(l[k], -hk, hk);
        if (q[k] != l[k]) inside = 0;
        double dk = hk - fabs(l[k]);
        if (dk < bestd) { bestd = dk; best = k; }
    }
    if (inside) {
        double hk = h[best] - rho;
        nout_local[0] = nout_local[1] = nout_local[2] = 0;
        nout_local[best] = l[best] >= 0 ? 1.0 : -1.0;
        q[best] = l[best] >= 0 ? hk : -hk;
    }
    return inside;
}
/* closest point of a cylinder (radius r, half height h, both shrunk by rho) to local point l */
static int cyl_closest(double *q, const double *l, double r, double h, double rho, double *nout_local) {
    double rr = r - rho, hh = h - rho, rad = sqrt(l[0] * l[0] + l[1] * l[1]);
    int in_r = rad <= rr, in_z = fabs(l[2]) <= hh;
    double sc = (rad > rr && rad > 1e-12) ? rr / rad : 1.0;
    q[0] = l[0] * sc; q[1] = l[1] * sc; q[2] = clampd(l[2], -hh, hh);
    if (in_r && in_z) {
        double dr = rr - rad, dz = hh - fabs(l[2]);
        nout_local[0] = nout_local[1] = nout_local[2] = 0;
        if (dz < dr || rad < 1e-12) { nout_local[2] = l[2] >= 0 ? 1.0 : -1.0; q[2] = l[2] >= 0 ? hh : -hh; }
        else { nout_local[0] = l[0] / rad; nout_local[1] = l[1] / rad; q[0] = l[0] / rad * rr; q[1] = l[1] / rad * rr; }
        return 1;
    }
    return 0;
}
/* closest point on segment (centre c, unit axis a, half length h) to p */
static void seg_closest(double *q, const double *c, const double *a, double h, const double *p) {
    double d[3] = {p[0] - c[0], p[1] - c[1], p[2] - c[2]};
    double t = clampd(dot3(d, a), -h, h);
    for (int k = 0; k < 3; k++) q[k] = c[k] + t * a[k];
}

/* generic core projection: closest point of geom core (shrunk by rho) to world point p */
static int core_closest(double *qw, const cgeom *g, double rho, const double *p, double *nout_w) {
    double l[3], q[3], nl[3] = {0, 0, 0};
    int inside = 0;
    if (g->type == 6) { to_local(l, g, p); inside = box_closest(q, l, g->size, rho, nl); }
    else if (g->type == 5) { to_local(l, g, p); inside = cyl_closest(q, l, g->size[0], g->size[1], rho, nl); }
    else if (g->type == 3) { double a[3]; colk(a, g->R, 2); seg_closest(qw, g->c, a, g->size[1], p); return 0; }
    else { memcpy(qw, g->c, sizeof(double) * 3); return 0; }
    to_world(qw, g, q);
    if (inside) for (int k = 0; k < 3; k++) nout_w[k] = g->R[3 * k] * nl[0] + g->R[3 * k + 1] * nl[1] + g->R[3 * k + 2] * nl[2];
    return inside;
}
static double core_radius(const cgeom *g, double rho) {
    if (g->type == 2 || g->type == 3) return g->size[0];
    return rho;
}
static double core_rho(const cgeom *g, int against_curved) {
    if (g->type == 5) { double m = g->size[0] < g->size[1] ? g->size[0] : g->size[1]; return 0.005 < 0.5 * m ? 0.005 : 0.5 * m; }
    if (g->type == 6 && against_curved) {
        double m = g->size[0] < g->size[1] ? g->size[0] : g->size[1];
        m = m < g->size[2] ? m : g->size[2];
        return 0.005 < 0.5 * m ? 0.005 : 0.5 * m;
    }
    return 0.0;
}

/* one contact between two convex geoms by alternating projections of their cores */
static int convex_pocs(cpoint *cp, const cgeom *g1, const cgeom *g2, double margin) {
    int curved = (g1->type == 5 || g2->type == 5);
    double rho1 = (g1->type == 2 || g1->type == 3) ? 0.0 : core_rho(g1, curved && g1->type == 6 ? 1 : (g1->type == 5));
    double rho2 = (g2->type == 2 || g2->type == 3) ? 0.0 : core_rho(g2, curved && g2->type == 6 ? 1 : (g2->type == 5));
    double r1 = core_radius(g1, rho1), r2 = core_radius(g2, rho2);
    double p1[3], p2[3], n1[3], n2[3];
    int in1 = 0, in2 = 0;
    memcpy(p1, g1->c, sizeof(p1));
    for (int it = 0; it < 16; it++) {
        double o1[3] = {p1[0], p1[1], p1[2]};
        in2 = core_closest(p2, g2, rho2, p1, n2);
        in1 = core_closest(p1, g1, rho1, p2, n1);
        if (it > 0 && o1[0] == p1[0] && o1[1] == p1[1] && o1[2] == p1[2]) break; /* exact fixed point: later sweeps repeat it */
    }
    in2 = core_closest(p2, g2, rho2, p1, n2);
    double d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, len = sqrt(dot3(d, d)), n[3], dist;
    if (len > 1e-9) {
        for (int k = 0; k < 3; k++) n[k] = d[k] / len;
        dist = len - r1 - r2;
    } else { /* cores touch or overlap: use the outward normal of whichever core reports containment */
        if (in2) { for (int k = 0; k < 3; k++) n[k] = -n2[k]; }
        else if (in1) { memcpy(n, n1, sizeof(n)); }
        else return 0;
        dist = -r1 - r2;
    }
    if (dist >= margin) return 0;
    cp->dist = dist;
    for (int k = 0; k < 3; k++) { cp->n[k] = n[k]; cp->pos[k] = 0.5 * ((p1[k] + n[k] * r1) + (p2[k] - n[k] * r2)); }
    return 1;
}

/* box-box: SAT + face clipping / edge-edge */
static int box_box_contacts(cpoint *out, const cgeom *g1, const cgeom *g2, double margin) {
    const double *c1 = g1->c, *c2 = g2->c, *R1 = g1->R, *R2 = g2->R, *h1 = g1->size, *h2 = g2->size;
    double d[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]}, T[3], Rm[9], A[9];
    for (int k = 0; k < 3; k++) T[k] = R1[k] * d[0] + R1[3 + k] * d[1] + R1[6 + k] * d[2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            Rm[3 * i + j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
            A[3 * i + j] = fabs(Rm[3 * i + j]);
        }
    double best = -1e30;
    int code = -1; /* 0..2 face of box1, 3..5 face of box2, 6.. edge i*3+j */
    for (int i = 0; i < 3; i++) {
        double rb = A[3 * i] * h2[0] + A[3 * i + 1] * h2[1] + A[3 * i + 2] * h2[2];
        double s = fabs(T[i]) - h1[i] - rb;
        if (s > best) { best = s; code = i; }
    }
    for (int j = 0; j < 3; j++) {
        double ra = A[j] * h1[0] + A[3 + j] * h1[1] + A[6 + j] * h1[2];
        double tp = T[0] * Rm[j] + T[1] * Rm[3 + j] + T[2] * Rm[6 + j];
        double s = fabs(tp) - ra - h2[j];
        if (s > best) { best = s; code = 3 + j; }
    }
    if (best >= margin) return 0;   /* separated on a face axis: no axis can bring the maximum below margin */
    double ebest = -1e30;
    int ecode = -1;
    double en[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; j++) {
            int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            double l2 = 1.0 - Rm[3 * i + j] * Rm[3 * i + j];
            if (l2 < 1e-6) continue;
            double ra = h1[i1] * A[3 * i2 + j] + h1[i2] * A[3 * i1 + j];
            double rb = h2[j1] * A[3 * i + j2] + h2[j2] * A[3 * i + j1];
            double tp = T[i2] * Rm[3 * i1 + j] - T[i1] * Rm[3 * i2 + j];
            double s = (fabs(tp) - ra - rb) / sqrt(l2);
            if (s > ebest) { ebest = s; ecode = 3 * i + j; }
        }
    }
    /* an edge axis must beat the best face axis by a clear margin to be preferred */
    int use_edge = (ecode >= 0 && ebest > best + 1e-6 + 0.05 * fabs(best));
    if ((use_edge ? ebest : best) >= margin) return 0;
    if (use_edge) {
        int i = ecode / 3, j = ecode % 3;
        double a[3], b[3];
        colk(a, R1, i); colk(b, R2, j);
        cross3(en, a, b);
        double l = sqrt(dot3(en, en));
        for (int k = 0; k < 3; k++) en[k] /= l;
        if (dot3(en, d) < 0) for (int k = 0; k < 3; k++) en[k] = -en[k];
        /* supporting edge of each box in direction +-en */
        double p1[3] = {c1[0], c1[1], c1[2]}, p2[3] = {c2[0], c2[1], c2[2]};
        for (int k = 0; k < 3; k++) {
            if (k != i) { double ax[3]; colk(ax, R1, k); double sg = dot3(ax, en) > 0 ? 1.0 : -1.0; for (int c = 0; c < 3; c++) p1[c] += sg * h1[k] * ax[c]; }
            if (k != j) { double ax[3]; colk(ax, R2, k); double sg = dot3(ax, en) > 0 ? -1.0 : 1.0; for (int c = 0; c < 3; c++) p2[c] += sg * h2[k] * ax[c]; }
        }
        /* closest points of the two lines p1 + s a, p2 + t b */
        double w[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, ab = dot3(a, b), aw = dot3(a, w), bw = dot3(b, w);
        double den = 1.0 - ab * ab, s = (ab * bw - aw) / den, t = (bw - ab * aw) / den;
        s = clampd(s, -h1[i], h1[i]); t = clampd(t, -h2[j], h2[j]);
        out->dist = ebest;
        for (int k = 0; k < 3; k++) { out->n[k] = en[k]; out->pos[k] = 0.5 * ((p1[k] + s * a[k]) + (p2[k] + t * b[k])); }
        return 1;
    }
    /* face contact: reference box owns the axis, incident box is clipped against it */
    const cgeom *gr = code < 3 ? g1 : g2, *gi = code < 3 ? g2 : g1;
    int ax = code < 3 ? code : code - 3;
    double n[3], dd[3] = {gi->c[0] - gr->c[0], gi->c[1] - gr->c[1], gi->c[2] - gr->c[2]};
    colk(n, gr->R, ax);
    if (dot3(n, dd) < 0) for (int k = 0; k < 3; k++) n[k] = -n[k]; /* from reference towards incident */
    /* incident face: the face of gi whose normal is most opposed to n */
    int iax = 0;
    double mind = 1e30, isg = 1;
    for (int k = 0; k < 3; k++) {
        double a[3];
        colk(a, gi->R, k);
        double dn = dot3(a, n);
        if (-fabs(dn) < mind) { mind = -fabs(dn); iax = k; isg = dn > 0 ? -1.0 : 1.0; }
    }
    int u = (iax + 1) % 3, v = (iax + 2) % 3;
    double poly[16][3], tmp[16][3];
    int np = 4;
    {
        double fa[3], ua[3], va[3];
        colk(fa, gi->R, iax); colk(ua, gi->R, u); colk(va, gi->R, v);
        const double su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
        for (int q = 0; q < 4; q++)
            for (int k = 0; k < 3; k++)
                poly[q][k] = gi->c[k] + isg * gi->size[iax] * fa[k] + su[q] * gi->size[u] * ua[k] + sv[q] * gi->size[v] * va[k];
    }
    /* clip against the four side planes of the reference face */
    int ru = (ax + 1) % 3, rv = (ax + 2) % 3;
    for (int side = 0; side < 4 && np > 0; side++) {
        double pa[3];
        colk(pa, gr->R, side < 2 ? ru : rv);
        double sg = (side % 2) ? -1.0 : 1.0, lim = gr->size[side < 2 ? ru : rv];
        int nn = 0;
        for (int q = 0; q < np; q++) {
            const double *P = poly[q], *Q = poly[(q + 1) % np];
            double dp = sg * ((P[0] - gr->c[0]) * pa[0] + (P[1] - gr->c[1]) * pa[1] + (P[2] - gr->c[2]) * pa[2]) - lim;
            double dq = sg * ((Q[0] - gr->c[0]) * pa[0] + (Q[1] - gr->c[1]) * pa[1] + (Q[2] - gr->c[2]) * pa[2]) - lim;
            if (dp <= 0) { memcpy(tmp[nn++], P, sizeof(double) * 3); }
            if ((dp <= 0) != (dq <= 0)) {
                double t = dp / (dp - dq);
                for (int k = 0; k < 3; k++) tmp[nn][k] = P[k] + t * (Q[k] - P[k]);
                nn++;
            }
        }
        np = nn;
        memcpy(poly, tmp, sizeof(double) * 3 * np);
    }
    /* depth of each surviving vertex below the reference face; keep the (at most) 4 deepest */
    double depth[16];
    int keep[16], nk = 0;
    for (int q = 0; q < np; q++) {
        depth[q] = (poly[q][0] - gr->c[0]) * n[0] + (poly[q][1] - gr->c[1]) * n[1] + (poly[q][2] - gr->c[2]) * n[2] - gr->size[ax];
        if (depth[q] < margin) keep[nk++] = q;
    }
    while (nk > 4) { /* drop the shallowest */
        int w = 0;
        for (int q = 1; q < nk; q++) if (depth[keep[q]] > depth[keep[w]]) w = q;
        for (int q = w; q < nk - 1; q++) keep[q] = keep[q + 1];
        nk--;
    }
    double flip = (gr == g1) ? 1.0 : -1.0; /* n points reference -> incident; contacts report geom1 -> geom2 */
    for (int q = 0; q < nk; q++) {
        const double *P = poly[keep[q]];
        out[q].dist = depth[keep[q]];
        for (int k = 0; k < 3; k++) { out[q].n[k] = flip * n[k]; out[q].pos[k] = P[k] - n[k] * 0.5 * depth[keep[q]]; }
    }
    return nk;
}

/* contacts of one pair; returns the number written (<= 4) */
static int pair_contacts(cpoint *out, const cgeom *ga, const cgeom *gb, double margin) {
    const cgeom *g1 = ga, *g2 = gb;
    int swapped = 0;
    if (g1->type > g2->type) { const cgeom *t = g1; g1 = g2; g2 = t; swapped = 1; }
    int n = 0, t1 = g1->type, t2 = g2->type;
    if (t1 == 0) { /* plane */
        double pn[3];
        colk(pn, g1->R, 2);
        if (t2 == 2 || t2 == 3) {
            double a[3] = {0, 0, 0};
            int ne = 1;
            if (t2 == 3) { colk(a, g2->R, 2); ne = 2; }
            for (int e = 0; e < ne; e++) {
                double sgn = (t2 == 3) ? (e ? -1.0 : 1.0) * g2->size[1] : 0.0, c[3];
                for (int k = 0; k < 3; k++) c[k] = g2->c[k] + sgn * a[k];
                double dist = (c[0] - g1->c[0]) * pn[0] + (c[1] - g1->c[1]) * pn[1] + (c[2] - g1->c[2]) * pn[2] - g2->size[0];
                if (dist >= margin) continue;
                out[n].dist = dist;
                for (int k = 0; k < 3; k++) { out[n].n[k] = pn[k]; out[n].pos[k] = c[k] - pn[k] * (g2->size[0] + 0.5 * dist); }
                n++;
            }
        } else if (t2 == 5) {
            /* plane - cylinder as MuJoCo's mjc_PlaneCylinder builds it: deepest rim point of the cap facing the plane, the rim
               point of the other cap in the same radial direction, two more rim points of the near cap (triangle) */
            double a[3], vec[3], v1[3], P[4][3];
            colk(a, g2->R, 2);
            double prj = dot3(pn, a);
            if (prj > 0) { for (int k = 0; k < 3; k++) a[k] = -a[k]; prj = -prj; }
            for (int k = 0; k < 3; k++) vec[k] = a[k] * prj - pn[k];
            double len = sqrt(dot3(vec, vec));
            if (len < 1e-12) { colk(vec, g2->R, 0); len = 1.0; }
            for (int k = 0; k < 3; k++) vec[k] *= g2->size[0] / len;
            const double hh = g2->size[1];
            cross3(v1, vec, a);
            for (int k = 0; k < 3; k++) {
                P[0][k] = g2->c[k] + hh * a[k] + vec[k];
                P[1][k] = g2->c[k] - hh * a[k] + vec[k];
                P[2][k] = g2->c[k] + hh * a[k] - 0.5 * vec[k] + 0.8660254037844386 * v1[k];
                P[3][k] = g2->c[k] + hh * a[k] - 0.5 * vec[k] - 0.8660254037844386 * v1[k];
            }
            for (int q = 0; q < 4; q++) {
                double dist = (P[q][0] - g1->c[0]) * pn[0] + (P[q][1] - g1->c[1]) * pn[1] + (P[q][2] - g1->c[2]) * pn[2];
                if (dist >= margin) continue;
                out[n].dist = dist;
                for (int k = 0; k < 3; k++) { out[n].n[k] = pn[k]; out[n].pos[k] = P[q][k] - pn[k] * 0.5 * dist; }
                n++;
            }
        } else if (t2 == 6) {
            for (int q = 0; q < 8 && n < 4; q++) {
                double l[3] = {(q & 1 ? 1 : -1) * g2->size[0], (q & 2 ? 1 : -1) * g2->size[1], (q & 4 ? 1 : -1) * g2->size[2]}, c[3];
                to_world(c, g2, l);
                double dist = (c[0] - g1->c[0]) * pn[0] + (c[1] - g1->c[1]) * pn[1] + (c[2] - g1->c[2]) * pn[2];
                if (dist >= margin) continue;
                out[n].dist = dist;
                for (int k = 0; k < 3; k++) { out[n].n[k] = pn[k]; out[n].pos[k] = c[k] - pn[k] * 0.5 * dist; }
                n++;
            }
        }
    } else if (t1 == 2) { /* sphere vs X */
        double q[3], nout[3] = {0, 0, 0};
        int inside = 0;
        if (t2 == 2) memcpy(q, g2->c, sizeof(q));
        else inside = core_closest(q, g2, 0.0, g1->c, nout);
        double r2 = (t2 == 2 || t2 == 3) ? g2->size[0] : 0.0;
        /* treat as sphere of radius r1 + r2 against the core point q */
        n = sphere_vs_point(out, g1->c, g1->size[0] + r2, q, inside, nout, margin, 1);
        if (n) {
            /* move the point from the core surface to midway between the two real surfaces */
            for (int k = 0; k < 3; k++) out->pos[k] = q[k] - out->n[k] * (r2 + 0.5 * out->dist);
        }
    } else if (t1 == 3 && t2 == 3) { /* capsule-capsule: segment-segment closest points */
        double a1[3], a2[3];
        colk(a1, g1->R, 2); colk(a2, g2->R, 2);
        double r[3] = {g1->c[0] - g2->c[0], g1->c[1] - g2->c[1], g1->c[2] - g2->c[2]};
        double b = dot3(a1, a2), c = dot3(a1, r), f = dot3(a2, r), den = 1.0 - b * b, s, t;
        s = den > 1e-9 ? clampd((b * f - c) / den, -g1->size[1], g1->size[1]) : 0.0;
        t = b * s + f;
        if (t < -g2->size[1]) { t = -g2->size[1]; s = clampd(b * t - c, -g1->size[1], g1->size[1]); }
        else if (t > g2->size[1]) { t = g2->size[1]; s = clampd(b * t - c, -g1->size[1], g1->size[1]); }
        double p1[3], p2[3];
        for (int k = 0; k < 3; k++) { p1[k] = g1->c[k] + s * a1[k]; p2[k] = g2->c[k] + t * a2[k]; }
        n = sphere_vs_point(out, p1, g1->size[0] + g2->size[0], p2, 0, NULL, margin, 1);
        if (n) for (int k = 0; k < 3; k++) out->pos[k] = p2[k] - out->n[k] * (g2->size[0] + 0.5 * out->dist);
    } else if (t1 == 6 && t2 == 6) {
        n = box_box_contacts(out, g1, g2, margin);
    } else if (t1 >= 3 && t2 >= 5) {
        n = convex_pocs(out, g1, g2, margin);
    }
    if (swapped) for (int q = 0; q < n; q++) for (int k = 0; k < 3; k++) out[q].n[k] = -out[q].n[k];
    return n;
}

/* velocity of world point p due to unit rate of dof k */
static void point_jac(double *j, const sv6 *S, const double *p) {
    double t[3];
    cross3(t, S->w, p);
    for (int k = 0; k < 3; k++) j[k] = S->v[k] + t[k];
}

int orc_contact_rows(const dyn_model *m, const dyn_data *D, const sv6 *S, crow *rows, int maxrows) {
    int nrow = 0;
    double gc[DMAXG][3], gR[DMAXG][9];
    for (int g = 0; g < m->ngeom; g++) {
        double Rl[9];
        int b = m->g_body[g];
        double w = m->g_quat[g][0], x = m->g_quat[g][1], y = m->g_quat[g][2], z = m->g_quat[g][3];
        Rl[0] = w * w + x * x - y * y - z * z; Rl[1] = 2 * (x * y - w * z); Rl[2] = 2 * (x * z + w * y);
        Rl[3] = 2 * (x * y + w * z); Rl[4] = w * w - x * x + y * y - z * z; Rl[5] = 2 * (y * z - w * x);
        Rl[6] = 2 * (x * z - w * y); Rl[7] = 2 * (y * z + w * x); Rl[8] = w * w - x * x - y * y + z * z;
        if (b < 0) { memcpy(gc[g], m->g_pos[g], sizeof(double) * 3); memcpy(gR[g], Rl, sizeof(Rl)); continue; }
        const double *X = D->xmat[b];
        for (int k = 0; k < 3; k++) gc[g][k] = D->xpos[b][k] + X[3 * k] * m->g_pos[g][0] + X[3 * k + 1] * m->g_pos[g][1] + X[3 * k + 2] * m->g_pos[g][2];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) gR[g][3 * r + c] = X[3 * r] * Rl[c] + X[3 * r + 1] * Rl[3 + c] + X[3 * r + 2] * Rl[6 + c];
    }
    for (int p = 0; p < m->npair; p++) {
        int a = m->p_g1[p], b = m->p_g2[p];
        double margin = m->g_margin[a] > m->g_margin[b] ? m->g_margin[a] : m->g_margin[b];
        if (m->g_type[a] != 0 && m->g_type[b] != 0) {
            double d[3] = {gc[b][0] - gc[a][0], gc[b][1] - gc[a][1], gc[b][2] - gc[a][2]}, bound = m->g_rbound[a] + m->g_rbound[b] + margin;
            if (dot3(d, d) > bound * bound) continue;
        }
        cgeom ga = {gc[a], gR[a], m->g_size[a], m->g_type[a]}, gb = {gc[b], gR[b], m->g_size[b], m->g_type[b]};
        cpoint cps[4];
        int nc = pair_contacts(cps, &ga, &gb, margin);
        for (int q = 0; q < nc; q++) {
            if (nrow + 3 > maxrows) return nrow;
            /* tangent frame */
            double *n = cps[q].n, t1[3], t2[3], ref[3] = {0, 0, 0};
            ref[fabs(n[0]) < 0.7 ? 0 : 1] = 1.0;
            cross3(t1, n, ref);
            double l = sqrt(dot3(t1, t1));
            for (int k = 0; k < 3; k++) t1[k] /= l;
            cross3(t2, n, t1);
            const double *dirs[3] = {n, t1, t2};
            double mu = m->g_friction[a][0] > m->g_friction[b][0] ? m->g_friction[a][0] : m->g_friction[b][0];
            for (int r = 0; r < 3; r++) {
                crow *row = &rows[nrow + r];
                memset(row, 0, sizeof(crow));
                row->type = r == 0 ? 1 : 2;
                row->pos = cps[q].dist; row->margin = margin; row->mu = mu;
                row->sig = p * 16 + q * 4 + r;
                for (int k = 0; k < 2; k++) row->solref[k] = 0.5 * (m->g_solref[a][k] + m->g_solref[b][k]);
                for (int k = 0; k < 5; k++) row->solimp[k] = 0.5 * (m->g_solimp[a][k] + m->g_solimp[b][k]);
                for (int side = 0; side < 2; side++) {
                    int body = side ? m->g_body[b] : m->g_body[a];
                    double sg = side ? 1.0 : -1.0;
                    while (body >= 0 && m->b_jtype[body] < 0) body = m->b_parent[body];
                    if (body < 0) continue;
                    int k = m->b_dadr[body] + (m->b_jtype[body] == 0 ? 5 : 0);
                    for (; k >= 0; k = m->d_parent[k]) {
                        double j[3];
                        point_jac(j, &S[k], cps[q].pos);
                        row->J[k] += sg * dot3(dirs[r], j);
                    }
                }
            }
            nrow += 3;
        }
    }
    return nrow;
}
