// Contact generation for the physics step (device + host, double precision).
//
// Replaces MuJoCo 2.0's collision functions + mj_makeConstraint for contacts inside mj_step
// (reference call site: env/base.py:392 sim.step()).  Candidate pairs come from the host
// (dynmodel.py applies MuJoCo's static filters); here: bounding-sphere cull with margin, contact
// points (position, normal geom1 -> geom2, signed distance), then one normal and two tangent
// constraint rows per point with Jacobians built from the joint motion axes.
//   closed form   plane-{sphere,capsule,box}, sphere-X, capsule-capsule
//   box-box       15-axis SAT, then face clipping (<= 4 points) or edge-edge closest points
//   other convex  closest points of the cores by alternating projections (16 sweeps); cylinders
//                 (and a box facing a cylinder) are a core shrunk by rho <= 5 mm swept by a sphere
// Pair parameters combine as in MuJoCo: margin = max, friction = max, solref / solimp = mean.
#pragma once
#include "dyn.cuh"

namespace mopa {

struct CPoint { double pos[3], n[3], dist; };
struct CGeom { const double *c, *R, *size; int type; };

DYN_HD inline void c_colk(double *a, const double *R, int k) { a[0] = R[k]; a[1] = R[3 + k]; a[2] = R[6 + k]; }
DYN_HD inline void c_to_local(double *l, const CGeom &g, const double *p) {
    double d[3] = {p[0] - g.c[0], p[1] - g.c[1], p[2] - g.c[2]};
    for (int k = 0; k < 3; k++) l[k] = g.R[k] * d[0] + g.R[3 + k] * d[1] + g.R[6 + k] * d[2];
}
DYN_HD inline void c_to_world(double *p, const CGeom &g, const double *l) {
    for (int k = 0; k < 3; k++) p[k] = g.c[k] + g.R[3 * k] * l[0] + g.R[3 * k + 1] * l[1] + g.R[3 * k + 2] * l[2];
}
DYN_HD inline double c_clamp(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

DYN_HD inline int sphere_vs_point(CPoint &cp, const double *s, double r, const double *q, int inside, const double *nout, double margin,
                                  int sphere_is_g1) {
    double d[3] = {s[0] - q[0], s[1] - q[1], s[2] - q[2]}, len = sqrt(d_dot(d, d)), n[3], dist;
    if (inside) { for (int k = 0; k < 3; k++) n[k] = nout[k]; dist = -len - r; }
    else {
        if (len < 1e-12) return 0;
        for (int k = 0; k < 3; k++) n[k] = d[k] / len;
        dist = len - r;
    }
    if (dist >= margin) return 0;
    cp.dist = dist;
    for (int k = 0; k < 3; k++) { cp.n[k] = sphere_is_g1 ? -n[k] : n[k]; cp.pos[k] = q[k] + n[k] * 0.5 * dist; }
    return 1;
}
DYN_HD inline int box_closest(double *q, const double *l, const double *h, double rho, double *nl) {
    int inside = 1, best = 0;
    double bestd = 1e30;
    for (int k = 0; k < 3; k++) {
        double hk = h[k] - rho;
        q[k] = c_clamp(l[k], -hk, hk);
        if (q[k] != l[k]) inside = 0;
        double dk = hk - fabs(l[k]);
        if (dk < bestd) { bestd = dk; best = k; }
    }
    if (inside) {
        double hk = h[best] - rho;
        nl[0] = nl[1] = nl[2] = 0;
        nl[best] = l[best] >= 0 ? 1.0 : -1.0;
        q[best] = l[best] >= 0 ? hk : -hk;
    }
    return inside;
}
DYN_HD inline int cyl_closest(double *q, const double *l, double r, double h, double rho, double *nl) {
    double rr = r - rho, hh = h - rho, rad = sqrt(l[0] * l[0] + l[1] * l[1]);
    int in_r = rad <= rr, in_z = fabs(l[2]) <= hh;
    double sc = (rad > rr && rad > 1e-12) ? rr / rad : 1.0;
    q[0] = l[0] * sc; q[1] = l[1] * sc; q[2] = c_clamp(l[2], -hh, hh);
    if (in_r && in_z) {
        double dr = rr - rad, dz = hh - fabs(l[2]);
        nl[0] = nl[1] = nl[2] = 0;
        if (dz < dr || rad < 1e-12) { nl[2] = l[2] >= 0 ? 1.0 : -1.0; q[2] = l[2] >= 0 ? hh : -hh; }
        else { nl[0] = l[0] / rad; nl[1] = l[1] / rad; q[0] = l[0] / rad * rr; q[1] = l[1] / rad * rr; }
        return 1;
    }
    return 0;
}
DYN_HD inline int core_closest(double *qw, const CGeom &g, double rho, const double *p, double *nout_w) {
    double l[3], q[3], nl[3] = {0, 0, 0};
    int inside = 0;
    if (g.type == 6) { c_to_local(l, g, p); inside = box_closest(q, l, g.size, rho, nl); }
    else if (g.type == 5) { c_to_local(l, g, p); inside = cyl_closest(q, l, g.size[0], g.size[1], rho, nl); }
    else if (g.type == 3) {
        double a[3], d[3] = {p[0] - g.c[0], p[1] - g.c[1], p[2] - g.c[2]};
        c_colk(a, g.R, 2);
        double t = c_clamp(d_dot(d, a), -g.size[1], g.size[1]);
        for (int k = 0; k < 3; k++) qw[k] = g.c[k] + t * a[k];
        return 0;
    } else { for (int k = 0; k < 3; k++) qw[k] = g.c[k]; return 0; }
    c_to_world(qw, g, q);
    if (inside) for (int k = 0; k < 3; k++) nout_w[k] = g.R[3 * k] * nl[0] + g.R[3 * k + 1] * nl[1] + g.R[3 * k + 2] * nl[2];
    return inside;
}
DYN_HD inline double core_rho(const CGeom &g, int against_curved) {
    if (g.type == 5) { double m = g.size[0] < g.size[1] ? g.size[0] : g.size[1]; return 0.005 < 0.5 * m ? 0.005 : 0.5 * m; }
    if (g.type == 6 && against_curved) {
        double m = g.size[0] < g.size[1] ? g.size[0] : g.size[1];
        m = m < g.size[2] ? m : g.size[2];
        return 0.005 < 0.5 * m ? 0.005 : 0.5 * m;
    }
    return 0.0;
}
DYN_HD inline int convex_pocs(CPoint &cp, const CGeom &g1, const CGeom &g2, double margin) {
    const int curved = (g1.type == 5 || g2.type == 5);
    const double rho1 = (g1.type == 2 || g1.type == 3) ? 0.0 : core_rho(g1, curved && g1.type == 6 ? 1 : (g1.type == 5));
    const double rho2 = (g2.type == 2 || g2.type == 3) ? 0.0 : core_rho(g2, curved && g2.type == 6 ? 1 : (g2.type == 5));
    const double r1 = (g1.type == 2 || g1.type == 3) ? g1.size[0] : rho1, r2 = (g2.type == 2 || g2.type == 3) ? g2.size[0] : rho2;
    double p1[3], p2[3], n1[3] = {0, 0, 0}, n2[3] = {0, 0, 0};
    int in1 = 0, in2 = 0;
    for (int k = 0; k < 3; k++) p1[k] = g1.c[k];
    for (int it = 0; it < 16; it++) {
        double o1[3] = {p1[0], p1[1], p1[2]};
        in2 = core_closest(p2, g2, rho2, p1, n2);
        in1 = core_closest(p1, g1, rho1, p2, n1);
        if (it > 0 && o1[0] == p1[0] && o1[1] == p1[1] && o1[2] == p1[2]) break;   // exact fixed point: later sweeps repeat it
    }
    in2 = core_closest(p2, g2, rho2, p1, n2);
    double d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, len = sqrt(d_dot(d, d)), n[3], dist;
    if (len > 1e-9) {
        for (int k = 0; k < 3; k++) n[k] = d[k] / len;
        dist = len - r1 - r2;
    } else {
        if (in2) { for (int k = 0; k < 3; k++) n[k] = -n2[k]; }
        else if (in1) { for (int k = 0; k < 3; k++) n[k] = n1[k]; }
        else return 0;
        dist = -r1 - r2;
    }
    if (dist >= margin) return 0;
    cp.dist = dist;
    for (int k = 0; k < 3; k++) { cp.n[k] = n[k]; cp.pos[k] = 0.5 * ((p1[k] + n[k] * r1) + (p2[k] - n[k] * r2)); }
    return 1;
}

// box-box, part 1: 15-axis SAT.  Returns 0 = separated by more than margin, 1 = edge-edge contact
// (ecode = 3 * i + j, depth ebest), 2 = face contact (code = face axis 0..5, see box_box_face).
DYN_HD inline int box_box_sat(const CGeom &g1, const CGeom &g2, double margin, int &code_out, double &depth_out) {
    const double *c1 = g1.c, *c2 = g2.c, *R1 = g1.R, *R2 = g2.R, *h1 = g1.size, *h2 = g2.size;
    double d[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]}, T[3], Rm[9], A[9];
    for (int k = 0; k < 3; k++) T[k] = R1[k] * d[0] + R1[3 + k] * d[1] + R1[6 + k] * d[2];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            Rm[3 * i + j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j];
            A[3 * i + j] = fabs(Rm[3 * i + j]);
        }
    double best = -1e30;
    int code = -1;
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
    for (int i = 0; i < 3; i++) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        for (int j = 0; j < 3; j++) {
            const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            double l2 = 1.0 - Rm[3 * i + j] * Rm[3 * i + j];
            if (l2 < 1e-6) continue;
            double ra = h1[i1] * A[3 * i2 + j] + h1[i2] * A[3 * i1 + j];
            double rb = h2[j1] * A[3 * i + j2] + h2[j2] * A[3 * i + j1];
            double tp = T[i2] * Rm[3 * i1 + j] - T[i1] * Rm[3 * i2 + j];
            double s = (fabs(tp) - ra - rb) / sqrt(l2);
            if (s > ebest) { ebest = s; ecode = 3 * i + j; }
        }
    }
    const int use_edge = (ecode >= 0 && ebest > best + 1e-6 + 0.05 * fabs(best));
    if ((use_edge ? ebest : best) >= margin) return 0;
    code_out = use_edge ? ecode : code;
    depth_out = use_edge ? ebest : best;
    return use_edge ? 1 : 2;
}
// part 2a: closest points of the two edges named by ecode
DYN_HD inline int box_box_edge(CPoint *out, const CGeom &g1, const CGeom &g2, int ecode, double ebest) {
    const double *c1 = g1.c, *c2 = g2.c, *R1 = g1.R, *R2 = g2.R, *h1 = g1.size, *h2 = g2.size;
    const double d[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
    const int i = ecode / 3, j = ecode % 3;
    double a[3], b[3], en[3];
    c_colk(a, R1, i); c_colk(b, R2, j);
    d_cross(en, a, b);
    double l = sqrt(d_dot(en, en));
    for (int k = 0; k < 3; k++) en[k] /= l;
    if (d_dot(en, d) < 0) for (int k = 0; k < 3; k++) en[k] = -en[k];
    double p1[3] = {c1[0], c1[1], c1[2]}, p2[3] = {c2[0], c2[1], c2[2]};
    for (int k = 0; k < 3; k++) {
        if (k != i) { double ax[3]; c_colk(ax, R1, k); double sg = d_dot(ax, en) > 0 ? 1.0 : -1.0; for (int c = 0; c < 3; c++) p1[c] += sg * h1[k] * ax[c]; }
        if (k != j) { double ax[3]; c_colk(ax, R2, k); double sg = d_dot(ax, en) > 0 ? -1.0 : 1.0; for (int c = 0; c < 3; c++) p2[c] += sg * h2[k] * ax[c]; }
    }
    double w[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, ab = d_dot(a, b), aw = d_dot(a, w), bw = d_dot(b, w);
    double den = 1.0 - ab * ab, s = (ab * bw - aw) / den, t = (bw - ab * aw) / den;
    s = c_clamp(s, -h1[i], h1[i]); t = c_clamp(t, -h2[j], h2[j]);
    out[0].dist = ebest;
    for (int k = 0; k < 3; k++) { out[0].n[k] = en[k]; out[0].pos[k] = 0.5 * ((p1[k] + s * a[k]) + (p2[k] + t * b[k])); }
    return 1;
}
// part 2b: face contact - the incident face of the other box clipped against the reference face
// (Sutherland-Hodgman; a quad clipped by four half planes has at most 8 vertices), <= 4 deepest points kept.
// poly / tmp [>= 8][3], depth [>= 8], keep [>= 8]: caller-provided scratch (shared memory in the warp kernel).
DYN_HD inline int box_box_face(CPoint *out, const CGeom &g1, const CGeom &g2, int code, double margin, double (*poly)[3], double (*tmp)[3],
                               double *depth, int *keep) {
    const CGeom &gr = code < 3 ? g1 : g2, &gi = code < 3 ? g2 : g1;
    const int ax = code < 3 ? code : code - 3;
    double n[3], dd[3] = {gi.c[0] - gr.c[0], gi.c[1] - gr.c[1], gi.c[2] - gr.c[2]};
    c_colk(n, gr.R, ax);
    if (d_dot(n, dd) < 0) for (int k = 0; k < 3; k++) n[k] = -n[k];
    int iax = 0;
    double mind = 1e30, isg = 1;
    for (int k = 0; k < 3; k++) {
        double a[3];
        c_colk(a, gi.R, k);
        double dn = d_dot(a, n);
        if (-fabs(dn) < mind) { mind = -fabs(dn); iax = k; isg = dn > 0 ? -1.0 : 1.0; }
    }
    const int u = (iax + 1) % 3, v = (iax + 2) % 3;
    int np = 4;
    {
        double fa[3], ua[3], va[3];
        c_colk(fa, gi.R, iax); c_colk(ua, gi.R, u); c_colk(va, gi.R, v);
        const double su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
        for (int q = 0; q < 4; q++)
            for (int k = 0; k < 3; k++)
                poly[q][k] = gi.c[k] + isg * gi.size[iax] * fa[k] + su[q] * gi.size[u] * ua[k] + sv[q] * gi.size[v] * va[k];
    }
    const int ru = (ax + 1) % 3, rv = (ax + 2) % 3;
    for (int side = 0; side < 4 && np > 0; side++) {
        double pa[3];
        c_colk(pa, gr.R, side < 2 ? ru : rv);
        const double sg = (side % 2) ? -1.0 : 1.0, lim = gr.size[side < 2 ? ru : rv];
        int nn = 0;
        for (int q = 0; q < np; q++) {
            const double *P = poly[q], *Q = poly[(q + 1) % np];
            double dp = sg * ((P[0] - gr.c[0]) * pa[0] + (P[1] - gr.c[1]) * pa[1] + (P[2] - gr.c[2]) * pa[2]) - lim;
            double dq = sg * ((Q[0] - gr.c[0]) * pa[0] + (Q[1] - gr.c[1]) * pa[1] + (Q[2] - gr.c[2]) * pa[2]) - lim;
            if (dp <= 0) { for (int k = 0; k < 3; k++) tmp[nn][k] = P[k]; nn++; }
            if ((dp <= 0) != (dq <= 0)) {
                double t = dp / (dp - dq);
                for (int k = 0; k < 3; k++) tmp[nn][k] = P[k] + t * (Q[k] - P[k]);
                nn++;
            }
        }
        np = nn;
        for (int q = 0; q < np; q++) for (int k = 0; k < 3; k++) poly[q][k] = tmp[q][k];
    }
    int nk = 0;
    for (int q = 0; q < np; q++) {
        depth[q] = (poly[q][0] - gr.c[0]) * n[0] + (poly[q][1] - gr.c[1]) * n[1] + (poly[q][2] - gr.c[2]) * n[2] - gr.size[ax];
        if (depth[q] < margin) keep[nk++] = q;
    }
    while (nk > 4) {
        int w = 0;
        for (int q = 1; q < nk; q++) if (depth[keep[q]] > depth[keep[w]]) w = q;
        for (int q = w; q < nk - 1; q++) keep[q] = keep[q + 1];
        nk--;
    }
    const double flip = (&gr == &g1) ? 1.0 : -1.0;
    for (int q = 0; q < nk; q++) {
        const double *P = poly[keep[q]];
        out[q].dist = depth[keep[q]];
        for (int k = 0; k < 3; k++) { out[q].n[k] = flip * n[k]; out[q].pos[k] = P[k] - n[k] * 0.5 * depth[keep[q]]; }
    }
    return nk;
}
DYN_HD inline int box_box_contacts(CPoint *out, const CGeom &g1, const CGeom &g2, double margin) {
    int code = 0;
    double depth = 0;
    const int kind = box_box_sat(g1, g2, margin, code, depth);
    if (kind == 0) return 0;
    if (kind == 1) return box_box_edge(out, g1, g2, code, depth);
    double poly[8][3], tmp[8][3], dep[8];
    int keep[8];
    return box_box_face(out, g1, g2, code, margin, poly, tmp, dep, keep);
}

DYN_HD inline int pair_contacts(CPoint *out, const CGeom &ga, const CGeom &gb, double margin) {
    const bool swapped = ga.type > gb.type;
    const CGeom &g1 = swapped ? gb : ga, &g2 = swapped ? ga : gb;
    int n = 0;
    const int t1 = g1.type, t2 = g2.type;
    if (t1 == 0) {
        double pn[3];
        c_colk(pn, g1.R, 2);
        if (t2 == 2 || t2 == 3) {
            double a[3] = {0, 0, 0};
            int ne = 1;
            if (t2 == 3) { c_colk(a, g2.R, 2); ne = 2; }
            for (int e = 0; e < ne; e++) {
                double sgn = (t2 == 3) ? (e ? -1.0 : 1.0) * g2.size[1] : 0.0, c[3];
                for (int k = 0; k < 3; k++) c[k] = g2.c[k] + sgn * a[k];
                double dist = (c[0] - g1.c[0]) * pn[0] + (c[1] - g1.c[1]) * pn[1] + (c[2] - g1.c[2]) * pn[2] - g2.size[0];
                if (dist >= margin) continue;
                out[n].dist = dist;
                for (int k = 0; k < 3; k++) { out[n].n[k] = pn[k]; out[n].pos[k] = c[k] - pn[k] * (g2.size[0] + 0.5 * dist); }
                n++;
            }
        } else if (t2 == 5) {
            // plane - cylinder, the construction of MuJoCo's mjc_PlaneCylinder: the deepest rim point of the cap that faces the
            // plane, the rim point of the other cap in the same radial direction, and - for a cylinder lying nearly flat on that
            // cap - two more points of its rim (a triangle with the first one).
            double a[3], vec[3];
            c_colk(a, g2.R, 2);
            double prj = d_dot(pn, a);
            if (prj > 0) { for (int k = 0; k < 3; k++) a[k] = -a[k]; prj = -prj; }   // a points towards the plane
            for (int k = 0; k < 3; k++) vec[k] = a[k] * prj - pn[k];                     // radial direction of steepest descent
            double len = sqrt(d_dot(vec, vec));
            if (len < 1e-12) {   // axis parallel to the normal: any radial direction (MuJoCo takes the cylinder's x axis)
                c_colk(vec, g2.R, 0);
                len = 1.0;
            }
            for (int k = 0; k < 3; k++) vec[k] *= g2.size[0] / len;
            const double hh = g2.size[1];
            double P[4][3], v1[3];
            d_cross(v1, vec, a);
            for (int k = 0; k < 3; k++) {
                P[0][k] = g2.c[k] + hh * a[k] + vec[k];
                P[1][k] = g2.c[k] - hh * a[k] + vec[k];
                P[2][k] = g2.c[k] + hh * a[k] - 0.5 * vec[k] + 0.8660254037844386 * v1[k];
                P[3][k] = g2.c[k] + hh * a[k] - 0.5 * vec[k] - 0.8660254037844386 * v1[k];
            }
            for (int q = 0; q < 4; q++) {
                const double dist = (P[q][0] - g1.c[0]) * pn[0] + (P[q][1] - g1.c[1]) * pn[1] + (P[q][2] - g1.c[2]) * pn[2];
                if (dist >= margin) continue;
                out[n].dist = dist;
                for (int k = 0; k < 3; k++) { out[n].n[k] = pn[k]; out[n].pos[k] = P[q][k] - pn[k] * 0.5 * dist; }
                n++;
            }
        } else if (t2 == 6) {
            for (int q = 0; q < 8 && n < 4; q++) {
                double l[3] = {(q & 1 ? 1 : -1) * g2.size[0], (q & 2 ? 1 : -1) * g2.size[1], (q & 4 ? 1 : -1) * g2.size[2]}, c[3];
                c_to_world(c, g2, l);
                double dist = (c[0] - g1.c[0]) * pn[0] + (c[1] - g1.c[1]) * pn[1] + (c[2] - g1.c[2]) * pn[2];
                if (dist >= margin) continue;
                out[n].dist = dist;
                for (int k = 0; k < 3; k++) { out[n].n[k] = pn[k]; out[n].pos[k] = c[k] - pn[k] * 0.5 * dist; }
                n++;
            }
        }
    } else if (t1 == 2) {
        double q[3], nout[3] = {0, 0, 0};
        int inside = 0;
        if (t2 == 2) { for (int k = 0; k < 3; k++) q[k] = g2.c[k]; }
        else inside = core_closest(q, g2, 0.0, g1.c, nout);
        const double r2 = (t2 == 2 || t2 == 3) ? g2.size[0] : 0.0;
        n = sphere_vs_point(out[0], g1.c, g1.size[0] + r2, q, inside, nout, margin, 1);
        if (n) for (int k = 0; k < 3; k++) out[0].pos[k] = q[k] - out[0].n[k] * (r2 + 0.5 * out[0].dist);
    } else if (t1 == 3 && t2 == 3) {
        double a1[3], a2[3];
        c_colk(a1, g1.R, 2); c_colk(a2, g2.R, 2);
        double r[3] = {g1.c[0] - g2.c[0], g1.c[1] - g2.c[1], g1.c[2] - g2.c[2]};
        double b = d_dot(a1, a2), c = d_dot(a1, r), f = d_dot(a2, r), den = 1.0 - b * b, s, t;
        s = den > 1e-9 ? c_clamp((b * f - c) / den, -g1.size[1], g1.size[1]) : 0.0;
        t = b * s + f;
        if (t < -g2.size[1]) { t = -g2.size[1]; s = c_clamp(b * t - c, -g1.size[1], g1.size[1]); }
        else if (t > g2.size[1]) { t = g2.size[1]; s = c_clamp(b * t - c, -g1.size[1], g1.size[1]); }
        double p1[3], p2[3];
        for (int k = 0; k < 3; k++) { p1[k] = g1.c[k] + s * a1[k]; p2[k] = g2.c[k] + t * a2[k]; }
        n = sphere_vs_point(out[0], p1, g1.size[0] + g2.size[0], p2, 0, nullptr, margin, 1);
        if (n) for (int k = 0; k < 3; k++) out[0].pos[k] = p2[k] - out[0].n[k] * (g2.size[0] + 0.5 * out[0].dist);
    } else if (t1 == 6 && t2 == 6) {
        n = box_box_contacts(out, g1, g2, margin);
    } else if (t1 >= 3 && t2 >= 5) {
        n = convex_pocs(out[0], g1, g2, margin);
    }
    if (swapped) for (int q = 0; q < n; q++) for (int k = 0; k < 3; k++) out[q].n[k] = -out[q].n[k];
    return n;
}

}  // namespace mopa
