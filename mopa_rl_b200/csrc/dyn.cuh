// Physics step (one mj_step equivalent) for one environment, double precision, device + host.
//
// Replaces the MuJoCo call inside BaseEnv._do_simulation (env/base.py:388-392: data.ctrl[:] = a;
// sim.step()) for the simulated sub-trees of the scene (include/mopa_dyn_desc.h): kinematics,
// composite-rigid-body inertia + armature, RNE bias, damping / actuator / applied forces, soft
// joint-limit and contact constraints solved by projected Gauss-Seidel on the dual, and
// semi-implicit Euler with implicit joint damping.  One thread integrates one environment; all
// 75 substeps of an env.step run on chip and only the state row goes back to HBM.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DYN_HD __host__ __device__
#else
#define DYN_HD
#endif

namespace mopa {

constexpr int DMAXB = 20;  // simulated bodies
constexpr int DMAXD = 16;  // simulated dofs
constexpr int DMAXA = 12;  // actuators
constexpr int DMAXG = 48;  // contact geoms
constexpr int DMAXC = 36;  // constraint rows
#define DYN_MINVAL 1e-15

struct DynDev {
    int nq, nv, nb, nd, nact, ngeom, npair, iterations;
    double h, g[3], tolerance;
    int b_parent[DMAXB], b_jtype[DMAXB], b_qadr[DMAXB], b_vadr[DMAXB], b_dadr[DMAXB];
    double b_pos[DMAXB][3], b_quat[DMAXB][4], b_rootpos[DMAXB][3], b_rootquat[DMAXB][4], b_jaxis[DMAXB][3], b_jpos[DMAXB][3];
    double b_qpos0[DMAXB], b_mass[DMAXB], b_ipos[DMAXB][3], b_iquat[DMAXB][4], b_inertia[DMAXB][3];
    int d_body[DMAXD], d_qadr[DMAXD], d_vadr[DMAXD], d_limited[DMAXD], d_parent[DMAXD];
    double d_armature[DMAXD], d_damping[DMAXD], d_range[DMAXD][2], d_solref[DMAXD][2], d_solimp[DMAXD][5], d_margin[DMAXD];
    int a_dof[DMAXA], a_kind[DMAXA], a_ctrllimited[DMAXA], a_forcelimited[DMAXA];
    double a_kp[DMAXA], a_kv[DMAXA], a_gear[DMAXA], a_ctrlrange[DMAXA][2], a_forcerange[DMAXA][2];
    int enable_contacts;
    int g_body[DMAXG], g_type[DMAXG];
    double g_pos[DMAXG][3], g_quat[DMAXG][4], g_size[DMAXG][3], g_rbound[DMAXG], g_margin[DMAXG], g_friction[DMAXG][3];
    double g_solref[DMAXG][2], g_solimp[DMAXG][5];
    int p_g1[512], p_g2[512];
};

struct Sv6 { double w[3], v[3]; };
struct SInert { double m, h[3], I[9]; };

struct DynData {  // what mjData would hold after the step (kinematics of the step's start state)
    double xpos[DMAXB][3], xquat[DMAXB][4], xmat[DMAXB][9];
    double bias[DMAXD];
    int ncon;
};

struct CRow {
    double J[DMAXD];
    double pos, margin, solref[2], solimp[5], mu;
    int type;  // 0 limit, 1 contact normal, 2 tangent
    int sig;   // identity of the row across substeps (warm start)
};
struct WarmStart { int n; int sig[DMAXC]; double f[DMAXC]; };  // constraint forces of the previous substep

DYN_HD inline void d_q2m(double *M, const double *q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
    M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
    M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
DYN_HD inline void d_qmul(double *r, const double *a, const double *b) {
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
DYN_HD inline void d_mv(double *r, const double *M, const double *v) {
    double a = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], b = M[3] * v[0] + M[4] * v[1] + M[5] * v[2],
           c = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    r[0] = a; r[1] = b; r[2] = c;
}
DYN_HD inline void d_cross(double *r, const double *a, const double *b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
DYN_HD inline double d_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
DYN_HD inline void sv_cross_motion(Sv6 &r, const Sv6 &a, const Sv6 &s) {
    double t1[3], t2[3];
    d_cross(r.w, a.w, s.w);
    d_cross(t1, a.w, s.v); d_cross(t2, a.v, s.w);
    for (int k = 0; k < 3; k++) r.v[k] = t1[k] + t2[k];
}
DYN_HD inline void sv_cross_force(Sv6 &r, const Sv6 &a, const Sv6 &f) {
    double t1[3], t2[3];
    d_cross(t1, a.w, f.w); d_cross(t2, a.v, f.v);
    for (int k = 0; k < 3; k++) r.w[k] = t1[k] + t2[k];
    d_cross(r.v, a.w, f.v);
}
DYN_HD inline void inert_apply(Sv6 &f, const SInert &I, const Sv6 &a) {
    double t[3], u[3];
    d_mv(t, I.I, a.w); d_cross(u, I.h, a.v);
    for (int k = 0; k < 3; k++) f.w[k] = t[k] + u[k];
    d_cross(u, I.h, a.w);
    for (int k = 0; k < 3; k++) f.v[k] = I.m * a.v[k] - u[k];
}
DYN_HD inline double sv_dot(const Sv6 &s, const Sv6 &f) { return d_dot(s.w, f.w) + d_dot(s.v, f.v); }

// impedance / reference parameters of a constraint row (mj_makeImpedance semantics)
DYN_HD inline void d_kbi(const DynDev &m, const double *solref, const double *solimp, double pos, double margin, double &K, double &B, double &imp) {
    double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
    double x = fabs(pos - margin) / (width > DYN_MINVAL ? width : DYN_MINVAL);
    double y;
    if (x >= 1) y = 1;
    else if (power <= 1) y = x;
    else if (x <= mid) y = pow(x, power) / pow(mid > DYN_MINVAL ? mid : DYN_MINVAL, power - 1);
    else y = 1 - pow(1 - x, power) / pow((1 - mid) > DYN_MINVAL ? (1 - mid) : DYN_MINVAL, power - 1);
    double im = dmin + y * (dmax - dmin);
    if (im < 1e-4) im = 1e-4;
    if (im > 0.9999) im = 0.9999;
    double tc = solref[0] > 2 * m.h ? solref[0] : 2 * m.h, dr = solref[1];
    double kd = dmax * dmax * tc * tc * dr * dr, bd = dmax * tc;
    K = 1 / (kd > DYN_MINVAL ? kd : DYN_MINVAL);
    B = 2 / (bd > DYN_MINVAL ? bd : DYN_MINVAL);
    imp = im;
}

struct DynDev; DYN_HD inline int contact_rows(const DynDev &m, const DynData &D, const Sv6 *S, CRow *rows, int maxrows);

// in-place Cholesky solve with the lower factor L (row-major DMAXD x DMAXD)
DYN_HD inline void chol_solve(const double *L, int nd, double *x) {
    for (int i = 0; i < nd; i++) {
        double s = x[i];
        for (int k = 0; k < i; k++) s -= L[i * DMAXD + k] * x[k];
        x[i] = s / L[i * DMAXD + i];
    }
    for (int i = nd - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < nd; k++) s -= L[k * DMAXD + i] * x[k];
        x[i] = s / L[i * DMAXD + i];
    }
}
DYN_HD inline void chol_factor(const double *M, const double *diag_add, double hscale, int nd, double *L) {
    for (int i = 0; i < nd; i++)
        for (int j = 0; j <= i; j++) {
            double s = M[i * DMAXD + j] + (i == j && diag_add ? hscale * diag_add[i] : 0.0);
            for (int k = 0; k < j; k++) s -= L[i * DMAXD + k] * L[j * DMAXD + k];
            L[i * DMAXD + j] = (i == j) ? sqrt(s) : s / L[j * DMAXD + j];
        }
}

// One mj_step.  qpos / qvel are FULL rows (nq / nv) updated in place.  `integrate` = false turns
// the call into the kinematics + bias part of mj_forward.
DYN_HD inline void dyn_substep(const DynDev &m, double *qpos, double *qvel, const double *ctrl, const double *applied, DynData &D,
                               bool integrate, WarmStart *warm = nullptr) {
    const int nb = m.nb, nd = m.nd;
    Sv6 S[DMAXD], vel[DMAXB], frc[DMAXB];
    SInert I[DMAXB];
    double qd[DMAXD];
    for (int i = 0; i < nd; i++) qd[i] = qvel[m.d_vadr[i]];
    for (int i = 0; i < nb; i++) {
        double Pp[3], Pq[4], PM[9];
        int p = m.b_parent[i];
        if (p >= 0) {
            for (int k = 0; k < 3; k++) Pp[k] = D.xpos[p][k];
            for (int k = 0; k < 4; k++) Pq[k] = D.xquat[p][k];
            for (int k = 0; k < 9; k++) PM[k] = D.xmat[p][k];
        } else {
            for (int k = 0; k < 3; k++) Pp[k] = m.b_rootpos[i][k];
            for (int k = 0; k < 4; k++) Pq[k] = m.b_rootquat[i][k];
            d_q2m(PM, Pq);
        }
        double pos[3], quat[4], t[3], R[9];
        d_mv(t, PM, m.b_pos[i]);
        for (int k = 0; k < 3; k++) pos[k] = Pp[k] + t[k];
        d_qmul(quat, Pq, m.b_quat[i]);
        const int jt = m.b_jtype[i], da = m.b_dadr[i];
        if (jt == 3) {
            double anchor[3], ql[4], ax[3], qn[4];
            d_q2m(R, quat);
            d_mv(t, R, m.b_jpos[i]);
            for (int k = 0; k < 3; k++) anchor[k] = pos[k] + t[k];
            double ang = qpos[m.b_qadr[i]] - m.b_qpos0[i], sn = sin(0.5 * ang), cs = cos(0.5 * ang);
            ql[0] = cs; ql[1] = sn * m.b_jaxis[i][0]; ql[2] = sn * m.b_jaxis[i][1]; ql[3] = sn * m.b_jaxis[i][2];
            d_qmul(qn, quat, ql);
            for (int k = 0; k < 4; k++) quat[k] = qn[k];
            d_q2m(R, quat);
            d_mv(t, R, m.b_jpos[i]);
            for (int k = 0; k < 3; k++) pos[k] = anchor[k] - t[k];
            d_mv(ax, R, m.b_jaxis[i]);
            for (int k = 0; k < 3; k++) S[da].w[k] = ax[k];
            d_cross(S[da].v, anchor, ax);
        } else if (jt == 2) {
            double ax[3];
            d_q2m(R, quat);
            d_mv(ax, R, m.b_jaxis[i]);
            double dq = qpos[m.b_qadr[i]] - m.b_qpos0[i];
            for (int k = 0; k < 3; k++) { pos[k] += ax[k] * dq; S[da].w[k] = 0; S[da].v[k] = ax[k]; }
        } else if (jt == 0) {
            const int a = m.b_qadr[i];
            for (int k = 0; k < 3; k++) pos[k] = qpos[a + k];
            double n = sqrt(qpos[a + 3] * qpos[a + 3] + qpos[a + 4] * qpos[a + 4] + qpos[a + 5] * qpos[a + 5] + qpos[a + 6] * qpos[a + 6]);
            for (int k = 0; k < 4; k++) quat[k] = qpos[a + 3 + k] / n;
            d_q2m(R, quat);
            for (int k = 0; k < 3; k++) {
                for (int c = 0; c < 3; c++) { S[da + k].w[c] = 0; S[da + k].v[c] = 0; }
                S[da + k].v[k] = 1;
                double e[3] = {R[k], R[3 + k], R[6 + k]};
                for (int c = 0; c < 3; c++) S[da + 3 + k].w[c] = e[c];
                d_cross(S[da + 3 + k].v, pos, e);
            }
        }
        d_q2m(R, quat);
        for (int k = 0; k < 3; k++) D.xpos[i][k] = pos[k];
        for (int k = 0; k < 4; k++) D.xquat[i][k] = quat[k];
        for (int k = 0; k < 9; k++) D.xmat[i][k] = R[k];
        double c[3], Ri[9], Mi[9], Iw[9];
        d_mv(t, R, m.b_ipos[i]);
        for (int k = 0; k < 3; k++) c[k] = pos[k] + t[k];
        d_q2m(Mi, m.b_iquat[i]);
        for (int r = 0; r < 3; r++)
            for (int cc = 0; cc < 3; cc++) Ri[3 * r + cc] = R[3 * r] * Mi[cc] + R[3 * r + 1] * Mi[3 + cc] + R[3 * r + 2] * Mi[6 + cc];
        for (int r = 0; r < 3; r++)
            for (int cc = 0; cc < 3; cc++)
                Iw[3 * r + cc] = Ri[3 * r] * m.b_inertia[i][0] * Ri[3 * cc] + Ri[3 * r + 1] * m.b_inertia[i][1] * Ri[3 * cc + 1] +
                                 Ri[3 * r + 2] * m.b_inertia[i][2] * Ri[3 * cc + 2];
        const double ms = m.b_mass[i], cc2 = d_dot(c, c);
        I[i].m = ms;
        for (int k = 0; k < 3; k++) I[i].h[k] = ms * c[k];
        for (int r = 0; r < 3; r++)
            for (int cc = 0; cc < 3; cc++) I[i].I[3 * r + cc] = Iw[3 * r + cc] + ms * ((r == cc ? cc2 : 0) - c[r] * c[cc]);
        if (p >= 0) vel[i] = vel[p];
        else for (int k = 0; k < 3; k++) { vel[i].w[k] = 0; vel[i].v[k] = 0; }
        const int ndj = jt < 0 ? 0 : (jt == 0 ? 6 : 1);
        for (int k = 0; k < ndj; k++)
            for (int c3 = 0; c3 < 3; c3++) { vel[i].w[c3] += S[da + k].w[c3] * qd[da + k]; vel[i].v[c3] += S[da + k].v[c3] * qd[da + k]; }
    }
    // bias: RNE with zero joint acceleration (acc reuses the frc array slot by slot)
    {
        Sv6 acc[DMAXB];
        for (int i = 0; i < nb; i++) {
            const int p = m.b_parent[i], jt = m.b_jtype[i], da = m.b_dadr[i];
            if (p >= 0) acc[i] = acc[p];
            else for (int k = 0; k < 3; k++) { acc[i].w[k] = 0; acc[i].v[k] = -m.g[k]; }
            const int ndj = jt < 0 ? 0 : (jt == 0 ? 6 : 1);
            for (int k = 0; k < ndj; k++) {
                if (jt == 0 && k < 3) continue;
                Sv6 sd;
                sv_cross_motion(sd, vel[i], S[da + k]);
                for (int c3 = 0; c3 < 3; c3++) { acc[i].w[c3] += sd.w[c3] * qd[da + k]; acc[i].v[c3] += sd.v[c3] * qd[da + k]; }
            }
            Sv6 Ia, Iv, vIv;
            inert_apply(Ia, I[i], acc[i]);
            inert_apply(Iv, I[i], vel[i]);
            sv_cross_force(vIv, vel[i], Iv);
            for (int c3 = 0; c3 < 3; c3++) { frc[i].w[c3] = Ia.w[c3] + vIv.w[c3]; frc[i].v[c3] = Ia.v[c3] + vIv.v[c3]; }
        }
    }
    for (int i = nb - 1; i >= 0; i--) {
        const int p = m.b_parent[i];
        if (p >= 0) for (int c3 = 0; c3 < 3; c3++) { frc[p].w[c3] += frc[i].w[c3]; frc[p].v[c3] += frc[i].v[c3]; }
    }
    double bias[DMAXD];
    for (int k = 0; k < nd; k++) { bias[k] = sv_dot(S[k], frc[m.d_body[k]]); D.bias[k] = bias[k]; }
    if (!integrate) return;
    // joint-space inertia by composite rigid bodies (I becomes the composite inertia)
    for (int i = nb - 1; i >= 0; i--) {
        const int p = m.b_parent[i];
        if (p < 0) continue;
        I[p].m += I[i].m;
        for (int k = 0; k < 3; k++) I[p].h[k] += I[i].h[k];
        for (int k = 0; k < 9; k++) I[p].I[k] += I[i].I[k];
    }
    double M[DMAXD * DMAXD];
    for (int i = 0; i < nd * DMAXD; i++) M[i] = 0;
    for (int i = 0; i < nd; i++) {
        Sv6 F;
        inert_apply(F, I[m.d_body[i]], S[i]);
        for (int j = i; j >= 0; j = m.d_parent[j]) { M[i * DMAXD + j] = sv_dot(S[j], F); M[j * DMAXD + i] = M[i * DMAXD + j]; }
        M[i * DMAXD + i] += m.d_armature[i];
    }
    double tau[DMAXD];
    for (int k = 0; k < nd; k++) tau[k] = -m.d_damping[k] * qd[k] - bias[k] + applied[k];
    for (int a = 0; a < m.nact; a++) {
        const int k = m.a_dof[a];
        double c = ctrl[a];
        if (m.a_ctrllimited[a]) c = c < m.a_ctrlrange[a][0] ? m.a_ctrlrange[a][0] : (c > m.a_ctrlrange[a][1] ? m.a_ctrlrange[a][1] : c);
        double q = m.d_qadr[k] >= 0 ? qpos[m.d_qadr[k]] : 0.0, f;
        if (m.a_kind[a] == 1) f = m.a_kp[a] * c - m.a_kp[a] * (m.a_gear[a] * q);
        else if (m.a_kind[a] == 2) f = m.a_kv[a] * c - m.a_kv[a] * (m.a_gear[a] * qd[k]);
        else f = c;
        if (m.a_forcelimited[a]) f = f < m.a_forcerange[a][0] ? m.a_forcerange[a][0] : (f > m.a_forcerange[a][1] ? m.a_forcerange[a][1] : f);
        tau[k] += m.a_gear[a] * f;
    }
    double L[DMAXD * DMAXD];
    chol_factor(M, nullptr, 0.0, nd, L);
    double qacc0[DMAXD];
    for (int k = 0; k < nd; k++) qacc0[k] = tau[k];
    chol_solve(L, nd, qacc0);
    // constraints
    CRow rows[DMAXC];
    int nc = 0;
    for (int k = 0; k < nd && nc < DMAXC; k++) {
        if (!m.d_limited[k] || m.d_qadr[k] < 0) continue;
        const double q = qpos[m.d_qadr[k]];
        for (int side = 0; side < 2; side++) {
            const double dist = side == 0 ? q - m.d_range[k][0] : m.d_range[k][1] - q;
            if (dist >= m.d_margin[k] || nc >= DMAXC) continue;
            CRow &r = rows[nc++];
            for (int j = 0; j < nd; j++) r.J[j] = 0;
            r.J[k] = side == 0 ? 1.0 : -1.0;
            r.pos = dist; r.margin = m.d_margin[k]; r.type = 0; r.mu = 0; r.sig = -(2 * k + side + 1);
            r.solref[0] = m.d_solref[k][0]; r.solref[1] = m.d_solref[k][1];
            for (int j = 0; j < 5; j++) r.solimp[j] = m.d_solimp[k][j];
        }
    }
    D.ncon = 0;
    if (m.enable_contacts && m.npair > 0) {
        const int n0 = nc;
        nc += contact_rows(m, D, S, rows + nc, DMAXC - nc);
        D.ncon = (nc - n0) / 3;
    }
    double fc[DMAXD];
    for (int k = 0; k < nd; k++) fc[k] = 0;
    if (nc > 0) {
        double MiJ[DMAXC][DMAXD], A[DMAXC * DMAXC], b[DMAXC], Rg[DMAXC], f[DMAXC];
        for (int r = 0; r < nc; r++) {
            for (int k = 0; k < nd; k++) MiJ[r][k] = rows[r].J[k];
            chol_solve(L, nd, MiJ[r]);
            f[r] = 0;
        }
        if (warm && warm->n == nc) {
            bool same = true;
            for (int r = 0; r < nc; r++) if (warm->sig[r] != rows[r].sig) same = false;
            if (same) for (int r = 0; r < nc; r++) f[r] = warm->f[r];
        }
        for (int r = 0; r < nc; r++)
            for (int s = 0; s < nc; s++) {
                double a = 0;
                for (int k = 0; k < nd; k++) a += rows[r].J[k] * MiJ[s][k];
                A[r * nc + s] = a;
            }
        for (int r = 0; r < nc; r++) {
            double K, B, imp, jv = 0, ja = 0;
            for (int k = 0; k < nd; k++) { jv += rows[r].J[k] * qd[k]; ja += rows[r].J[k] * qacc0[k]; }
            d_kbi(m, rows[r].solref, rows[r].solimp, rows[r].pos, rows[r].margin, K, B, imp);
            const double aref = rows[r].type <= 1 ? (-B * jv - K * imp * (rows[r].pos - rows[r].margin)) : (-B * jv);
            Rg[r] = (1 - imp) / imp * A[r * nc + r];
            if (Rg[r] < DYN_MINVAL) Rg[r] = DYN_MINVAL;
            b[r] = ja - aref;
        }
        double trM = 0;
        for (int k = 0; k < nd; k++) trM += M[k * DMAXD + k];
        const double scale = 1.0 / (trM > DYN_MINVAL ? trM : DYN_MINVAL);
        for (int it = 0; it < m.iterations; it++) {
            double imp = 0;
            for (int r = 0; r < nc; r++) {
                if (rows[r].type >= 2) continue;
                double res = b[r] + Rg[r] * f[r];
                for (int s = 0; s < nc; s++) res += A[r * nc + s] * f[s];
                double fn = f[r] - res / (A[r * nc + r] + Rg[r]);
                fn = fn > 0 ? fn : 0;
                imp += 0.5 * (A[r * nc + r] + Rg[r]) * (fn - f[r]) * (fn - f[r]);
                f[r] = fn;
                if (rows[r].type == 1) {
                    for (int t = 1; t <= 2; t++) {
                        const int q = r + t;
                        double rs = b[q] + Rg[q] * f[q];
                        for (int s = 0; s < nc; s++) rs += A[q * nc + s] * f[s];
                        const double ft_new = f[q] - rs / (A[q * nc + q] + Rg[q]);
                        imp += 0.5 * (A[q * nc + q] + Rg[q]) * (ft_new - f[q]) * (ft_new - f[q]);
                        f[q] = ft_new;
                    }
                    const double lim = rows[r].mu * f[r], ft = sqrt(f[r + 1] * f[r + 1] + f[r + 2] * f[r + 2]);
                    if (ft > lim) {
                        const double sc = ft > DYN_MINVAL ? lim / ft : 0;
                        for (int t = 1; t <= 2; t++) {
                            const int q = r + t;
                            const double fs = f[q] * sc;
                            imp += 0.5 * (A[q * nc + q] + Rg[q]) * (fs - f[q]) * (fs - f[q]);
                            f[q] = fs;
                        }
                    }
                }
            }
            if (scale * imp < m.tolerance) break;
        }
        for (int r = 0; r < nc; r++)
            for (int k = 0; k < nd; k++) fc[k] += rows[r].J[k] * f[r];
        if (warm) { warm->n = nc; for (int r = 0; r < nc; r++) { warm->sig[r] = rows[r].sig; warm->f[r] = f[r]; } }
    } else if (warm)
        warm->n = 0;
    // semi-implicit Euler with implicit joint damping
    chol_factor(M, m.d_damping, m.h, nd, L);
    double rhs[DMAXD];
    for (int k = 0; k < nd; k++) rhs[k] = tau[k] + fc[k];
    chol_solve(L, nd, rhs);
    for (int k = 0; k < nd; k++) { qd[k] += m.h * rhs[k]; qvel[m.d_vadr[k]] = qd[k]; }
    for (int i = 0; i < nb; i++) {
        const int jt = m.b_jtype[i], da = m.b_dadr[i], a = m.b_qadr[i];
        if (jt == 2 || jt == 3) qpos[a] += m.h * qd[da];
        else if (jt == 0) {
            for (int k = 0; k < 3; k++) qpos[a + k] += m.h * qd[da + k];
            double w[3] = {qd[da + 3], qd[da + 4], qd[da + 5]}, n = sqrt(d_dot(w, w)), ang = n * m.h;
            if (ang > 0) {
                double sn = sin(0.5 * ang) / n, dq[4] = {cos(0.5 * ang), w[0] * sn, w[1] * sn, w[2] * sn}, qn[4];
                d_qmul(qn, qpos + a + 3, dq);
                double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
                for (int k = 0; k < 4; k++) qpos[a + 3 + k] = qn[k] / nn;
            }
        }
    }
}

}  // namespace mopa
