/* TEST INFRASTRUCTURE — CPU oracle for the MoPA-RL hot path.  Not part of the product.
 *
 * Physics-step oracle (double precision): restates what one `sim.step()` of the reference does
 * (env/base.py:388-392 -> MuJoCo 2.0 mj_step, closed binary, absent here -> PARITY UNPINNED
 * against MuJoCo itself) for the simulated sub-trees handed over in mopa_dyn_desc, following
 * SURVEY.md App. B.4:
 *   kinematics -> composite-rigid-body inertia M (+ armature) -> RNE bias (gravity, Coriolis)
 *   -> passive damping, position/velocity actuators with ctrl/force clamps, qfrc_applied
 *   -> qacc_smooth = M^-1 tau -> soft constraints (joint limits, frictional contacts with
 *   solref/solimp impedance and elliptic cones) -> semi-implicit Euler with implicit damping.
 * The constraint forces solve MuJoCo's primal convex problem with the Newton solver the reference
 * uses (elliptic cones, exact line search, warm start from the previous substep's acceleration,
 * `iterations` / `tolerance` from the XML).  Documented departures: the regulariser uses
 * the exact diagonal of J M^-1 J^T instead of MuJoCo's precomputed invweight0 approximation; contacts
 * are condim-3 (no torsional / rolling friction); no noslip post-pass.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/mopa_dyn_desc.h"

#include "orc_dyn.h"

/* ---------------------------------------------------------------- small math */
static void q2m(double *M, const double *q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
    M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
    M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
static void qmul(double *r, const double *a, const double *b) {
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void mv(double *r, const double *M, const double *v) {
    double a = M[0] * v[0] + M[1] * v[1] + M[2] * v[2], b = M[3] * v[0] + M[4] * v[1] + M[5] * v[2],
           c = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    r[0] = a; r[1] = b; r[2] = c;
}
static void cross(double *r, const double *a, const double *b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* spatial algebra about the world origin */
static void sv_cross_motion(sv6 *r, const sv6 *a, const sv6 *s) { /* a x^ s */
    double t1[3], t2[3];
    cross(r->w, a->w, s->w);
    cross(t1, a->w, s->v); cross(t2, a->v, s->w);
    for (int k = 0; k < 3; k++) r->v[k] = t1[k] + t2[k];
}
static void sv_cross_force(sv6 *r, const sv6 *a, const sv6 *f) { /* a x* f */
    double t1[3], t2[3];
    cross(t1, a->w, f->w); cross(t2, a->v, f->v);
    for (int k = 0; k < 3; k++) r->w[k] = t1[k] + t2[k];
    cross(r->v, a->w, f->v);
}
static void inert_apply(sv6 *f, const sinert *I, const sv6 *a) {
    double t[3], u[3];
    mv(t, I->I, a->w); cross(u, I->h, a->v);
    for (int k = 0; k < 3; k++) f->w[k] = t[k] + u[k];
    cross(u, I->h, a->w);
    for (int k = 0; k < 3; k++) f->v[k] = I->m * a->v[k] - u[k];
}
static double sv_dot(const sv6 *s, const sv6 *f) { return dot(s->w, f->w) + dot(s->v, f->v); }

/* ---------------------------------------------------------------- model */
void *orc_dyn_create(const mopa_dyn_desc *d) {
    if (d->nb > DMAXB || d->nd > DMAXD || d->nact > DMAXA || d->ngeom > DMAXG) return NULL;
    dyn_model *m = (dyn_model *)calloc(1, sizeof(dyn_model));
    m->nq = d->nq; m->nv = d->nv; m->nb = d->nb; m->nd = d->nd; m->nact = d->nact; m->ngeom = d->ngeom; m->npair = d->npair;
    m->iterations = d->iterations; m->h = d->timestep; m->tolerance = d->tolerance;
    memcpy(m->g, d->gravity, sizeof(m->g));
    for (int i = 0; i < d->nb; i++) {
        m->b_parent[i] = d->b_parent[i]; m->b_jtype[i] = d->b_jtype[i]; m->b_qadr[i] = d->b_qadr[i];
        m->b_vadr[i] = d->b_vadr[i]; m->b_dadr[i] = d->b_dadr[i]; m->b_qpos0[i] = d->b_qpos0[i]; m->b_mass[i] = d->b_mass[i];
        for (int k = 0; k < 3; k++) {
            m->b_pos[i][k] = d->b_pos[3 * i + k]; m->b_rootpos[i][k] = d->b_rootpos[3 * i + k];
            m->b_jaxis[i][k] = d->b_jaxis[3 * i + k]; m->b_jpos[i][k] = d->b_jpos[3 * i + k];
            m->b_ipos[i][k] = d->b_ipos[3 * i + k]; m->b_inertia[i][k] = d->b_inertia[3 * i + k];
        }
        for (int k = 0; k < 4; k++) {
            m->b_quat[i][k] = d->b_quat[4 * i + k]; m->b_rootquat[i][k] = d->b_rootquat[4 * i + k]; m->b_iquat[i][k] = d->b_iquat[4 * i + k];
        }
    }
    for (int i = 0; i < d->nd; i++) {
        m->d_body[i] = d->d_body[i]; m->d_qadr[i] = d->d_qadr[i]; m->d_vadr[i] = d->d_vadr[i]; m->d_limited[i] = d->d_limited[i];
        m->d_armature[i] = d->d_armature[i]; m->d_damping[i] = d->d_damping[i]; m->d_margin[i] = d->d_margin[i];
        for (int k = 0; k < 2; k++) { m->d_range[i][k] = d->d_range[2 * i + k]; m->d_solref[i][k] = d->d_solref[2 * i + k]; }
        for (int k = 0; k < 5; k++) m->d_solimp[i][k] = d->d_solimp[5 * i + k];
    }
    /* parent dof: previous dof of the same body, else last dof of the nearest ancestor that has dofs */
    for (int i = 0; i < d->nd; i++) {
        int b = m->d_body[i];
        if (i > 0 && m->d_body[i - 1] == b) { m->d_parent[i] = i - 1; continue; }
        int p = m->b_parent[b];
        while (p >= 0 && m->b_jtype[p] < 0) p = m->b_parent[p];
        if (p < 0) m->d_parent[i] = -1;
        else m->d_parent[i] = m->b_dadr[p] + (m->b_jtype[p] == 0 ? 5 : 0);
    }
    for (int i = 0; i < d->nact; i++) {
        m->a_dof[i] = d->a_dof[i]; m->a_kind[i] = d->a_kind[i]; m->a_ctrllimited[i] = d->a_ctrllimited[i];
        m->a_forcelimited[i] = d->a_forcelimited[i]; m->a_kp[i] = d->a_kp[i]; m->a_kv[i] = d->a_kv[i]; m->a_gear[i] = d->a_gear[i];
        for (int k = 0; k < 2; k++) { m->a_ctrlrange[i][k] = d->a_ctrlrange[2 * i + k]; m->a_forcerange[i][k] = d->a_forcerange[2 * i + k]; }
    }
    for (int i = 0; i < d->ngeom; i++) {
        m->g_body[i] = d->g_body[i]; m->g_type[i] = d->g_type[i]; m->g_condim[i] = d->g_condim[i];
        m->g_rbound[i] = d->g_rbound[i]; m->g_margin[i] = d->g_margin[i];
        for (int k = 0; k < 3; k++) { m->g_pos[i][k] = d->g_pos[3 * i + k]; m->g_size[i][k] = d->g_size[3 * i + k]; m->g_friction[i][k] = d->g_friction[3 * i + k]; }
        for (int k = 0; k < 4; k++) m->g_quat[i][k] = d->g_quat[4 * i + k];
        for (int k = 0; k < 2; k++) m->g_solref[i][k] = d->g_solref[2 * i + k];
        for (int k = 0; k < 5; k++) m->g_solimp[i][k] = d->g_solimp[5 * i + k];
    }
    m->p_g1 = (int *)malloc(sizeof(int) * (d->npair + 1));
    m->p_g2 = (int *)malloc(sizeof(int) * (d->npair + 1));
    for (int i = 0; i < d->npair; i++) { m->p_g1[i] = d->p_g1[i]; m->p_g2[i] = d->p_g2[i]; }
    m->enable_contacts = 1;
    m->max_rows = DMAXC;
    return m;
}
void orc_dyn_destroy(void *h) {
    dyn_model *m = (dyn_model *)h;
    if (!m) return;
    free(m->p_g1); free(m->p_g2); free(m);
}
static long g_pgs_calls = 0, g_pgs_sweeps = 0;
static int g_last_ncon = 0, g_last_g1[DMAXC / 3], g_last_g2[DMAXC / 3]; /* contact list (simulated-geom indices) of the same call */
static double g_last_cforce = 0; /* contact-force metric of the latest orc_dyn_step call (single-threaded test use) */
void orc_pgs_stats(long *out, int reset) { out[0] = g_pgs_calls; out[1] = g_pgs_sweeps; if (reset) g_pgs_calls = g_pgs_sweeps = 0; }
void orc_dyn_enable_contacts(void *h, int on) { ((dyn_model *)h)->enable_contacts = on; }
/* constraint-row capacity (the env kernel keeps 24 rows for small scenes, 32 for large ones; excess contacts are dropped in pair order) */
/* 0 = semi-implicit Euler (Sawyer scenes), 1 = RK4 (Pusher: <option integrator="RK4">) */
void orc_dyn_set_integrator(void *h, int rk4) { ((dyn_model *)h)->integrator = rk4 ? 1 : 0; }
void orc_dyn_set_max_rows(void *h, int n) { ((dyn_model *)h)->max_rows = n < 3 ? 3 : (n > DMAXC ? DMAXC : n); }

/* impedance / reference parameters of one constraint row (mj_makeImpedance semantics) */
static void kbi(const dyn_model *m, const double *solref, const double *solimp, double pos, double margin, double *K, double *B, double *imp) {
    double dmin = solimp[0], dmax = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
    double x = fabs(pos - margin) / (width > MINVAL ? width : MINVAL);
    double y;
    if (x >= 1) y = 1;
    else if (power <= 1) y = x;
    else if (x <= mid) y = pow(x, power) / pow(mid > MINVAL ? mid : MINVAL, power - 1);
    else y = 1 - pow(1 - x, power) / pow((1 - mid) > MINVAL ? (1 - mid) : MINVAL, power - 1);
    double im = dmin + y * (dmax - dmin);
    if (im < 1e-4) im = 1e-4;
    if (im > 0.9999) im = 0.9999;
    double tc = solref[0] > 2 * m->h ? solref[0] : 2 * m->h, dr = solref[1];
    double kd = dmax * dmax * tc * tc * dr * dr, bd = dmax * tc;
    *K = 1 / (kd > MINVAL ? kd : MINVAL);
    *B = 2 / (bd > MINVAL ? bd : MINVAL);
    *imp = im;
}


/* penalty of one elliptic contact at x = (x_n, x_t1, x_t2): returns s, writes grad[3], H[9] (row-major) and the zone
   (0 top: inactive, 1 bottom: quadratic, 2 middle: cone surface) */
static double orc_cone(double mu, double D, const double *x, double *grad, double *H, int *zone) {
    const double t = sqrt(x[1] * x[1] + x[2] * x[2]);
    for (int k = 0; k < 9; k++) H[k] = 0;
    grad[0] = grad[1] = grad[2] = 0;
    if (x[0] >= mu * t) { *zone = 0; return 0.0; }
    if (mu * x[0] + t <= 0) {
        *zone = 1;
        for (int k = 0; k < 3; k++) { grad[k] = D * x[k]; H[4 * k] = D; }
        return 0.5 * D * (x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    }
    *zone = 2;
    const double Dm = D / (1 + mu * mu), e = x[0] - mu * t, u[3] = {1.0, -mu * x[1] / t, -mu * x[2] / t};
    for (int p = 0; p < 3; p++) { grad[p] = Dm * e * u[p]; for (int q = 0; q < 3; q++) H[3 * p + q] = Dm * u[p] * u[q]; }
    const double c = -Dm * e * mu / t;   /* > 0: curvature of the cone surface */
    H[4] += c * (1 - x[1] * x[1] / (t * t)); H[5] += c * (-x[1] * x[2] / (t * t));
    H[7] += c * (-x[1] * x[2] / (t * t));    H[8] += c * (1 - x[2] * x[2] / (t * t));
    return 0.5 * Dm * e * e;
}

/* one mj_step.  qpos[nq], qvel[nv] updated in place; ctrl[nact]; applied[nd] = qfrc_applied on the
   simulated dofs; data receives the kinematics / bias of this step. */
/* mj_integratePos for the simulated joints: hinge / slide linear, free joints by the exponential map of the angular velocity */
static void integrate_pos(const dyn_model *m, double *qpos, const double *qd, double h) {
    for (int i = 0; i < m->nb; i++) {
        int jt = m->b_jtype[i], da = m->b_dadr[i], a = m->b_qadr[i];
        if (jt == 2 || jt == 3) qpos[a] += h * qd[da];
        else if (jt == 0) {
            for (int k = 0; k < 3; k++) qpos[a + k] += h * qd[da + k];
            double w[3] = {qd[da + 3], qd[da + 4], qd[da + 5]}, n = sqrt(dot(w, w)), ang = n * h;
            if (ang > 0) {
                double sn = sin(0.5 * ang) / n, dq[4] = {cos(0.5 * ang), w[0] * sn, w[1] * sn, w[2] * sn}, qn[4];
                qmul(qn, qpos + a + 3, dq);
                double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
                for (int k = 0; k < 4; k++) qpos[a + 3 + k] = qn[k] / nn;
            }
        }
    }
}

/* acc_out != NULL: forward dynamics only - write qacc = M^-1 (tau + J^T f) of the simulated dofs and leave the state alone
   (one stage of the Runge-Kutta integrator); NULL: one semi-implicit Euler mj_step. */
static void substep(const dyn_model *m, double *qpos, double *qvel, const double *ctrl, const double *applied, dyn_data *D, warm_t *warm,
                    double *acc_out) {
    const int nb = m->nb, nd = m->nd;
    sv6 S[DMAXD], vel[DMAXB], acc[DMAXB], frc[DMAXB];
    sinert I[DMAXB], Ic[DMAXB];
    double qd[DMAXD];
    for (int i = 0; i < nd; i++) qd[i] = qvel[m->d_vadr[i]];
    /* ---- kinematics */
    for (int i = 0; i < nb; i++) {
        const double *Pp, *Pq;
        double PM[9];
        int p = m->b_parent[i];
        if (p >= 0) { Pp = D->xpos[p]; Pq = D->xquat[p]; memcpy(PM, D->xmat[p], sizeof(PM)); }
        else { Pp = m->b_rootpos[i]; Pq = m->b_rootquat[i]; q2m(PM, Pq); }
        double pos[3], quat[4], t[3], R[9];
        mv(t, PM, m->b_pos[i]);
        for (int k = 0; k < 3; k++) pos[k] = Pp[k] + t[k];
        qmul(quat, Pq, m->b_quat[i]);
        int jt = m->b_jtype[i], da = m->b_dadr[i];
        if (jt == 3) { /* hinge */
            double anchor[3], ql[4], ax[3];
            q2m(R, quat);
            mv(t, R, m->b_jpos[i]);
            for (int k = 0; k < 3; k++) anchor[k] = pos[k] + t[k];
            double ang = qpos[m->b_qadr[i]] - m->b_qpos0[i], sn = sin(0.5 * ang), cs = cos(0.5 * ang);
            ql[0] = cs; ql[1] = sn * m->b_jaxis[i][0]; ql[2] = sn * m->b_jaxis[i][1]; ql[3] = sn * m->b_jaxis[i][2];
            double qn[4];
            qmul(qn, quat, ql);
            memcpy(quat, qn, sizeof(qn));
            q2m(R, quat);
            mv(t, R, m->b_jpos[i]);
            for (int k = 0; k < 3; k++) pos[k] = anchor[k] - t[k];
            mv(ax, R, m->b_jaxis[i]);
            memcpy(S[da].w, ax, sizeof(ax));
            cross(S[da].v, anchor, ax);
        } else if (jt == 2) { /* slide */
            double ax[3];
            q2m(R, quat);
            mv(ax, R, m->b_jaxis[i]);
            double dq = qpos[m->b_qadr[i]] - m->b_qpos0[i];
            for (int k = 0; k < 3; k++) pos[k] += ax[k] * dq;
            S[da].w[0] = S[da].w[1] = S[da].w[2] = 0;
            memcpy(S[da].v, ax, sizeof(ax));
        } else if (jt == 0) { /* free */
            int a = m->b_qadr[i];
            for (int k = 0; k < 3; k++) pos[k] = qpos[a + k];
            double n = sqrt(qpos[a + 3] * qpos[a + 3] + qpos[a + 4] * qpos[a + 4] + qpos[a + 5] * qpos[a + 5] + qpos[a + 6] * qpos[a + 6]);
            for (int k = 0; k < 4; k++) quat[k] = qpos[a + 3 + k] / n;
            q2m(R, quat);
            for (int k = 0; k < 3; k++) {
                S[da + k].w[0] = S[da + k].w[1] = S[da + k].w[2] = 0;
                S[da + k].v[0] = S[da + k].v[1] = S[da + k].v[2] = 0;
                S[da + k].v[k] = 1;
                double e[3] = {R[k], R[3 + k], R[6 + k]}; /* body axis k in world */
                memcpy(S[da + 3 + k].w, e, sizeof(e));
                cross(S[da + 3 + k].v, pos, e);
            }
        }
        q2m(R, quat);
        memcpy(D->xpos[i], pos, sizeof(pos)); memcpy(D->xquat[i], quat, sizeof(quat)); memcpy(D->xmat[i], R, sizeof(R));
        /* spatial inertia about the world origin */
        double c[3], Ri[9], Mi[9], Iw[9];
        mv(t, R, m->b_ipos[i]);
        for (int k = 0; k < 3; k++) c[k] = pos[k] + t[k];
        q2m(Mi, m->b_iquat[i]);
        for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) Ri[3 * r + cc] = R[3 * r] * Mi[cc] + R[3 * r + 1] * Mi[3 + cc] + R[3 * r + 2] * Mi[6 + cc];
        for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++)
            Iw[3 * r + cc] = Ri[3 * r] * m->b_inertia[i][0] * Ri[3 * cc] + Ri[3 * r + 1] * m->b_inertia[i][1] * Ri[3 * cc + 1] + Ri[3 * r + 2] * m->b_inertia[i][2] * Ri[3 * cc + 2];
        memcpy(D->com[i], c, sizeof(c));
        double ms = m->b_mass[i], cc2 = dot(c, c);
        I[i].m = ms;
        for (int k = 0; k < 3; k++) I[i].h[k] = ms * c[k];
        for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) I[i].I[3 * r + cc] = Iw[3 * r + cc] + ms * ((r == cc ? cc2 : 0) - c[r] * c[cc]);
        /* velocity */
        if (p >= 0) vel[i] = vel[p]; else memset(&vel[i], 0, sizeof(sv6));
        int ndj = jt < 0 ? 0 : (jt == 0 ? 6 : 1);
        for (int k = 0; k < ndj; k++)
            for (int c3 = 0; c3 < 3; c3++) { vel[i].w[c3] += S[da + k].w[c3] * qd[da + k]; vel[i].v[c3] += S[da + k].v[c3] * qd[da + k]; }
    }
    /* ---- bias forces: RNE with zero joint acceleration, gravity as base acceleration */
    for (int i = 0; i < nb; i++) {
        int p = m->b_parent[i], jt = m->b_jtype[i], da = m->b_dadr[i];
        if (p >= 0) acc[i] = acc[p];
        else { memset(&acc[i], 0, sizeof(sv6)); for (int k = 0; k < 3; k++) acc[i].v[k] = -m->g[k]; }
        int ndj = jt < 0 ? 0 : (jt == 0 ? 6 : 1);
        for (int k = 0; k < ndj; k++) {
            if (jt == 0 && k < 3) continue; /* world-fixed translation axes: dS/dt = 0 */
            sv6 sd;
            sv_cross_motion(&sd, &vel[i], &S[da + k]);
            for (int c3 = 0; c3 < 3; c3++) { acc[i].w[c3] += sd.w[c3] * qd[da + k]; acc[i].v[c3] += sd.v[c3] * qd[da + k]; }
        }
        sv6 Ia, Iv, vIv;
        inert_apply(&Ia, &I[i], &acc[i]);
        inert_apply(&Iv, &I[i], &vel[i]);
        sv_cross_force(&vIv, &vel[i], &Iv);
        for (int c3 = 0; c3 < 3; c3++) { frc[i].w[c3] = Ia.w[c3] + vIv.w[c3]; frc[i].v[c3] = Ia.v[c3] + vIv.v[c3]; }
    }
    for (int i = nb - 1; i >= 0; i--) {
        int p = m->b_parent[i];
        if (p >= 0) for (int c3 = 0; c3 < 3; c3++) { frc[p].w[c3] += frc[i].w[c3]; frc[p].v[c3] += frc[i].v[c3]; }
    }
    double bias[DMAXD];
    for (int k = 0; k < nd; k++) bias[k] = sv_dot(&S[k], &frc[m->d_body[k]]);
    /* ---- joint-space inertia: composite rigid bodies */
    for (int i = 0; i < nb; i++) Ic[i] = I[i];
    for (int i = nb - 1; i >= 0; i--) {
        int p = m->b_parent[i];
        if (p < 0) continue;
        Ic[p].m += Ic[i].m;
        for (int k = 0; k < 3; k++) Ic[p].h[k] += Ic[i].h[k];
        for (int k = 0; k < 9; k++) Ic[p].I[k] += Ic[i].I[k];
    }
    double M[DMAXD][DMAXD];
    memset(M, 0, sizeof(M));
    for (int i = 0; i < nd; i++) {
        sv6 F;
        inert_apply(&F, &Ic[m->d_body[i]], &S[i]);
        for (int j = i; j >= 0; j = m->d_parent[j]) { M[i][j] = sv_dot(&S[j], &F); M[j][i] = M[i][j]; }
        M[i][i] += m->d_armature[i];
    }
    for (int i = 0; i < nd; i++) for (int j = 0; j < nd; j++) D->M[i * nd + j] = M[i][j];
    /* ---- applied / passive / actuator forces */
    double tau[DMAXD];
    for (int k = 0; k < nd; k++) tau[k] = -m->d_damping[k] * qd[k] - bias[k] + applied[k];
    for (int a = 0; a < m->nact; a++) {
        int k = m->a_dof[a];
        double c = ctrl[a];
        if (m->a_ctrllimited[a]) c = c < m->a_ctrlrange[a][0] ? m->a_ctrlrange[a][0] : (c > m->a_ctrlrange[a][1] ? m->a_ctrlrange[a][1] : c);
        double q = m->d_qadr[k] >= 0 ? qpos[m->d_qadr[k]] : 0.0, f;
        if (m->a_kind[a] == 1) f = m->a_kp[a] * c - m->a_kp[a] * (m->a_gear[a] * q);
        else if (m->a_kind[a] == 2) f = m->a_kv[a] * c - m->a_kv[a] * (m->a_gear[a] * qd[k]);
        else f = c;
        if (m->a_forcelimited[a]) f = f < m->a_forcerange[a][0] ? m->a_forcerange[a][0] : (f > m->a_forcerange[a][1] ? m->a_forcerange[a][1] : f);
        tau[k] += m->a_gear[a] * f;
    }
    /* ---- Cholesky M = L L^T, qacc_smooth */
    double L[DMAXD][DMAXD];
    memset(L, 0, sizeof(L));
    for (int i = 0; i < nd; i++)
        for (int j = 0; j <= i; j++) {
            double s = M[i][j];
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            L[i][j] = (i == j) ? sqrt(s) : s / L[j][j];
        }
#define CHOL_SOLVE(Lm, x)                                                                    \
    do {                                                                                     \
        for (int i_ = 0; i_ < nd; i_++) { double s_ = (x)[i_]; for (int k_ = 0; k_ < i_; k_++) s_ -= Lm[i_][k_] * (x)[k_]; (x)[i_] = s_ / Lm[i_][i_]; } \
        for (int i_ = nd - 1; i_ >= 0; i_--) { double s_ = (x)[i_]; for (int k_ = i_ + 1; k_ < nd; k_++) s_ -= Lm[k_][i_] * (x)[k_]; (x)[i_] = s_ / Lm[i_][i_]; } \
    } while (0)
    double qacc0[DMAXD];
    memcpy(qacc0, tau, sizeof(double) * nd);
    CHOL_SOLVE(L, qacc0);
    /* ---- constraints */
    crow *rows = (crow *)malloc(sizeof(crow) * DMAXC);
    int nc = 0;
    const int maxrows = m->max_rows;
    for (int k = 0; k < nd && nc < maxrows; k++) {
        if (!m->d_limited[k] || m->d_qadr[k] < 0) continue;
        double q = qpos[m->d_qadr[k]];
        for (int side = 0; side < 2; side++) {
            double dist = side == 0 ? q - m->d_range[k][0] : m->d_range[k][1] - q;
            if (dist >= m->d_margin[k] || nc >= maxrows) continue;
            crow *r = &rows[nc++];
            memset(r, 0, sizeof(crow));
            r->J[k] = side == 0 ? 1.0 : -1.0;
            r->pos = dist; r->margin = m->d_margin[k]; r->type = 0; r->sig = -(2 * k + side + 1);
            memcpy(r->solref, m->d_solref[k], sizeof(r->solref)); memcpy(r->solimp, m->d_solimp[k], sizeof(r->solimp));
        }
    }
    D->ncon = 0;
    D->cforce = 0;
    if (m->enable_contacts && m->npair > 0) {
        int n0 = nc;
        nc += orc_contact_rows(m, D, S, rows + nc, maxrows - nc);
        D->ncon = (nc - n0) / 3;
        for (int c = 0; c < D->ncon && c < DMAXC / 3; c++) D->con_pair[c] = rows[n0 + 3 * c].sig / 16; /* sig = pair*16 + point*4 + dir */
    }
    double fc[DMAXD];
    memset(fc, 0, sizeof(fc));
    if (nc > 0) {
        /* ---- constraint forces: MuJoCo's primal problem
         *   min_a  1/2 (a - a0)^T M (a - a0) + sum_c s_c(J_c a - aref_c)
         * solved by Newton's method with an exact line search (mj_solNewton).  s is the quadratic penalty
         * 1/2 D x^2 for x < 0 on a joint-limit row and, for a frictional contact (rows n, t1, t2 sharing the
         * normal row's D, impratio 1), 1/2 D dist^2(x, K*) with K* = {x_n >= mu |x_t|} the dual of the elliptic
         * friction cone: zero inside K* ("top zone"), 1/2 D |x|^2 in the polar cone mu x_n + |x_t| <= 0
         * ("bottom zone"), 1/2 D (x_n - mu |x_t|)^2 / (1 + mu^2) in between ("middle zone") -
         * mj_constraintUpdate's elliptic branch.  Strictly convex, so the minimiser is unique. */
        double *aref = (double *)malloc(sizeof(double) * nc), *Dr = (double *)malloc(sizeof(double) * nc);
        for (int r = 0; r < nc; r++) {
            double K, B, imp, jv = 0, y[DMAXD];
            for (int k = 0; k < nd; k++) jv += rows[r].J[k] * qd[k];
            kbi(m, rows[r].solref, rows[r].solimp, rows[r].pos, rows[r].margin, &K, &B, &imp);
            aref[r] = rows[r].type <= 1 ? (-B * jv - K * imp * (rows[r].pos - rows[r].margin)) : (-B * jv);
            /* regulariser R = (1 - imp) / imp * (J M^-1 J^T)_rr ; y = L^-1 J^T */
            double diag = 0;
            for (int k = 0; k < nd; k++) {
                double sacc = rows[r].J[k];
                for (int j = 0; j < k; j++) sacc -= L[k][j] * y[j];
                y[k] = sacc / L[k][k];
                diag += y[k] * y[k];
            }
            double Rg = (1 - imp) / imp * diag;
            if (Rg < MINVAL) Rg = MINVAL;
            Dr[r] = 1.0 / Rg;
        }
        for (int r = 0; r < nc; r++) if (rows[r].type == 2) Dr[r] = Dr[r - (rows[r - 1].type == 1 ? 1 : 2)]; /* friction rows take the normal row's D */
        double trM = 0;
        for (int k = 0; k < nd; k++) trM += M[k][k];
        const double scale = 1.0 / (trM > MINVAL ? trM : MINVAL); /* 1 / (meaninertia * nv) */
        double a[DMAXD], g[DMAXD], H[DMAXD][DMAXD], gs[DMAXC], x[DMAXC], Hc[DMAXC][9], cost = 0;
        int zone[DMAXC];
#define NEWTON_EVAL(avec, want_hess)                                                                              \
    do {                                                                                                          \
        double mat_[DMAXD];                                                                                       \
        cost = 0;                                                                                                 \
        for (int i_ = 0; i_ < nd; i_++) { double s_ = -tau[i_]; for (int j_ = 0; j_ < nd; j_++) s_ += M[i_][j_] * (avec)[j_]; mat_[i_] = s_; } \
        for (int i_ = 0; i_ < nd; i_++) cost += 0.5 * ((avec)[i_] - qacc0[i_]) * mat_[i_];                        \
        for (int r_ = 0; r_ < nc; r_++) { double s_ = -aref[r_]; for (int k_ = 0; k_ < nd; k_++) s_ += rows[r_].J[k_] * (avec)[k_]; x[r_] = s_; } \
        for (int r_ = 0; r_ < nc; r_++) {                                                                         \
            if (rows[r_].type == 0) { zone[r_] = x[r_] < 0; gs[r_] = zone[r_] ? Dr[r_] * x[r_] : 0.0; if (zone[r_]) cost += 0.5 * Dr[r_] * x[r_] * x[r_]; } \
            else if (rows[r_].type == 1) cost += orc_cone(rows[r_].mu, Dr[r_], x + r_, gs + r_, Hc[r_], &zone[r_]); \
        }                                                                                                         \
        for (int i_ = 0; i_ < nd; i_++) { double s_ = mat_[i_]; for (int r_ = 0; r_ < nc; r_++) s_ += rows[r_].J[i_] * gs[r_]; g[i_] = s_; } \
        if (want_hess) {                                                                                          \
            for (int i_ = 0; i_ < nd; i_++) for (int j_ = 0; j_ <= i_; j_++) {                                    \
                double s_ = M[i_][j_];                                                                            \
                for (int r_ = 0; r_ < nc; r_++) {                                                                 \
                    if (rows[r_].type == 0) { if (zone[r_]) s_ += Dr[r_] * rows[r_].J[i_] * rows[r_].J[j_]; }     \
                    else if (rows[r_].type == 1 && zone[r_])                                                      \
                        for (int p_ = 0; p_ < 3; p_++) for (int q_ = 0; q_ < 3; q_++) s_ += rows[r_ + p_].J[i_] * Hc[r_][3 * p_ + q_] * rows[r_ + q_].J[j_]; \
                }                                                                                                 \
                H[i_][j_] = s_;                                                                                   \
            }                                                                                                     \
        }                                                                                                         \
    } while (0)
        /* warm start: the previous substep's acceleration unless the unconstrained one costs less */
        memcpy(a, qacc0, sizeof(double) * nd);
        NEWTON_EVAL(a, 0);
        if (warm && warm->have_a) {
            double c0 = cost;
            NEWTON_EVAL(warm->a, 0);
            if (cost < c0) memcpy(a, warm->a, sizeof(double) * nd);
        }
        NEWTON_EVAL(a, 1);
        g_pgs_calls++;
        for (int it = 0; it < m->iterations; it++) {
            double gn = 0;
            for (int k = 0; k < nd; k++) gn += g[k] * g[k];
            if (scale * sqrt(gn) < m->tolerance) break;
            g_pgs_sweeps++;
            /* search direction p = -H^-1 g */
            double Lh2[DMAXD][DMAXD], p[DMAXD], Mp[DMAXD], jp[DMAXC];
            for (int i = 0; i < nd; i++)
                for (int j = 0; j <= i; j++) {
                    double sacc = H[i][j];
                    for (int k = 0; k < j; k++) sacc -= Lh2[i][k] * Lh2[j][k];
                    Lh2[i][j] = (i == j) ? sqrt(sacc) : sacc / Lh2[j][j];
                }
            for (int k = 0; k < nd; k++) p[k] = -g[k];
            CHOL_SOLVE(Lh2, p);
            /* exact line search on phi(alpha) = cost(a + alpha p): safeguarded Newton iteration on phi' */
            double q1 = 0, q2 = 0;
            for (int i = 0; i < nd; i++) { double sacc = 0; for (int j = 0; j < nd; j++) sacc += M[i][j] * p[j]; Mp[i] = sacc; }
            for (int i = 0; i < nd; i++) { double ma = -tau[i]; for (int j = 0; j < nd; j++) ma += M[i][j] * a[j]; q1 += p[i] * ma; q2 += p[i] * Mp[i]; }
            for (int r = 0; r < nc; r++) { double sacc = 0; for (int k = 0; k < nd; k++) sacc += rows[r].J[k] * p[k]; jp[r] = sacc; }
            double alpha = 1.0, lo = 0.0, hi = -1.0, d0 = 0;
            for (int ls = 0; ls < 24; ls++) {
                double d1 = q1 + alpha * q2, d2 = q2;
                for (int r = 0; r < nc; r++) {
                    if (rows[r].type == 0) { double xa = x[r] + alpha * jp[r]; if (xa < 0) { d1 += Dr[r] * xa * jp[r]; d2 += Dr[r] * jp[r] * jp[r]; } }
                    else if (rows[r].type == 1) {
                        double xa[3] = {x[r] + alpha * jp[r], x[r + 1] + alpha * jp[r + 1], x[r + 2] + alpha * jp[r + 2]}, gg[3], hh[9];
                        int z;
                        orc_cone(rows[r].mu, Dr[r], xa, gg, hh, &z);
                        if (z) for (int p_ = 0; p_ < 3; p_++) { d1 += gg[p_] * jp[r + p_]; for (int q_ = 0; q_ < 3; q_++) d2 += jp[r + p_] * hh[3 * p_ + q_] * jp[r + q_]; }
                    }
                }
                if (ls == 0) {   /* derivative at alpha = 1 first; phi'(0) = g.p gives the tolerance */
                    for (int k = 0; k < nd; k++) d0 += g[k] * p[k];
                }
                if (fabs(d1) <= 1e-6 * fabs(d0)) break;
                if (d1 < 0) lo = alpha; else hi = alpha;
                double an = alpha - d1 / d2;
                if (hi >= 0) { if (!(an > lo && an < hi)) an = 0.5 * (lo + hi); }
                else if (!(an > lo)) an = 2 * alpha;
                alpha = an;
            }
            for (int k = 0; k < nd; k++) a[k] += alpha * p[k];
            double old = cost;
            NEWTON_EVAL(a, 1);
            if (scale * (old - cost) < m->tolerance) break;
        }
#undef NEWTON_EVAL
        for (int r = 0; r < nc; r++)
            for (int k = 0; k < nd; k++) fc[k] -= rows[r].J[k] * gs[r];   /* constraint force f = -grad s */
        for (int r = 0; r < nc; r++) if (rows[r].type >= 1) D->cforce += fabs(gs[r]);
        if (warm) { warm->have_a = 1; memcpy(warm->a, a, sizeof(double) * nd); }
        free(aref); free(Dr);
    }
    else if (warm) warm->have_a = 0;
    free(rows);
    if (acc_out) { /* explicit stage: qacc = M^-1 (tau + J^T f), damping is part of tau */
        for (int k = 0; k < nd; k++) acc_out[k] = tau[k] + fc[k];
        CHOL_SOLVE(L, acc_out);
        memcpy(D->bias, bias, sizeof(double) * nd);
        return;
    }
    /* ---- semi-implicit Euler with implicit joint damping: (M + h D) qacc = tau + J^T f */
    double Lh[DMAXD][DMAXD], rhs[DMAXD];
    memset(Lh, 0, sizeof(Lh));
    for (int i = 0; i < nd; i++)
        for (int j = 0; j <= i; j++) {
            double s = M[i][j] + (i == j ? m->h * m->d_damping[i] : 0.0);
            for (int k = 0; k < j; k++) s -= Lh[i][k] * Lh[j][k];
            Lh[i][j] = (i == j) ? sqrt(s) : s / Lh[j][j];
        }
    for (int k = 0; k < nd; k++) rhs[k] = tau[k] + fc[k];
    CHOL_SOLVE(Lh, rhs);
    for (int k = 0; k < nd; k++) { qd[k] += m->h * rhs[k]; qvel[m->d_vadr[k]] = qd[k]; }
    integrate_pos(m, qpos, qd, m->h);
    memcpy(D->bias, bias, sizeof(double) * nd);
}

/* One mj_step with the 4th-order Runge-Kutta integrator (mj_RungeKutta, N = 4: A = [[1/2], [0, 1/2], [0, 0, 1]],
   B = [1/6, 1/3, 1/3, 1/6]); every stage runs the full forward dynamics incl. collision and the constraint solver. */
static void rk4_step(const dyn_model *m, double *qpos, double *qvel, const double *ctrl, const double *applied, dyn_data *D, warm_t *warm) {
    static const double A[3][3] = {{0.5, 0, 0}, {0, 0.5, 0}, {0, 0, 1.0}}, B[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
    const int nd = m->nd;
    double q0[64], v0[64], Fv[4][DMAXD], Fa[4][DMAXD], dv[DMAXD], da[DMAXD];
    memcpy(q0, qpos, sizeof(double) * m->nq);
    memcpy(v0, qvel, sizeof(double) * m->nv);
    for (int i = 0; i < 4; i++) {
        if (i > 0) { /* X[i] = X[0] + h * sum_j A[i-1][j] F[j] */
            for (int k = 0; k < nd; k++) {
                dv[k] = da[k] = 0;
                for (int j = 0; j < i; j++) { dv[k] += A[i - 1][j] * Fv[j][k]; da[k] += A[i - 1][j] * Fa[j][k]; }
            }
            memcpy(qpos, q0, sizeof(double) * m->nq);
            memcpy(qvel, v0, sizeof(double) * m->nv);
            integrate_pos(m, qpos, dv, m->h);
            for (int k = 0; k < nd; k++) qvel[m->d_vadr[k]] = v0[m->d_vadr[k]] + m->h * da[k];
        }
        for (int k = 0; k < nd; k++) Fv[i][k] = qvel[m->d_vadr[k]];
        substep(m, qpos, qvel, ctrl, applied, D, warm, Fa[i]);   /* mjData keeps the frames / contacts of the LAST stage */
    }
    for (int k = 0; k < nd; k++) {
        dv[k] = da[k] = 0;
        for (int j = 0; j < 4; j++) { dv[k] += B[j] * Fv[j][k]; da[k] += B[j] * Fa[j][k]; }
    }
    memcpy(qpos, q0, sizeof(double) * m->nq);
    memcpy(qvel, v0, sizeof(double) * m->nv);
    integrate_pos(m, qpos, dv, m->h);
    for (int k = 0; k < nd; k++) qvel[m->d_vadr[k]] = v0[m->d_vadr[k]] + m->h * da[k];
}

/* n mj_step calls with constant ctrl.  comp[nd]: 1 where qfrc_applied tracks the previous step's
   qfrc_bias (the gravity compensation the Sawyer envs apply, sawyer_push_obstacle.py:188-202);
   bias_prev[nd] carried between calls.  xpos/xquat (nullable) receive the body frames mjData
   holds after the last step (computed at the start of that step). */
int orc_dyn_step(void *h, double *qpos, double *qvel, const double *ctrl, const int32_t *comp, double *bias_prev, int nsub,
                 double *xpos, double *xquat, int32_t *ncon) {
    dyn_model *m = (dyn_model *)h;
    dyn_data D;
    memset(&D, 0, sizeof(D));
    double applied[DMAXD];
    warm_t warm;
    warm.have_a = 0; /* the warm start lives for the substeps of one call (one env.step) */
    for (int s = 0; s < nsub; s++) {
        for (int k = 0; k < m->nd; k++) applied[k] = comp[k] ? bias_prev[k] : 0.0;
        if (m->integrator == 1) rk4_step(m, qpos, qvel, ctrl, applied, &D, &warm);
        else substep(m, qpos, qvel, ctrl, applied, &D, &warm, NULL);
        memcpy(bias_prev, D.bias, sizeof(double) * m->nd);
    }
    if (xpos) for (int i = 0; i < m->nb; i++) for (int k = 0; k < 3; k++) xpos[3 * i + k] = D.xpos[i][k];
    if (xquat) for (int i = 0; i < m->nb; i++) for (int k = 0; k < 4; k++) xquat[4 * i + k] = D.xquat[i][k];
    if (ncon) *ncon = D.ncon;
    g_last_cforce = D.cforce;
    g_last_ncon = D.ncon < DMAXC / 3 ? D.ncon : DMAXC / 3;
    for (int c = 0; c < g_last_ncon; c++) { g_last_g1[c] = m->p_g1[D.con_pair[c]]; g_last_g2[c] = m->p_g2[D.con_pair[c]]; }
    return 0;
}
double orc_dyn_last_contact_force(void) { return g_last_cforce; }
/* mjData.contact[i].geom1 / geom2 after the step (sim.data.ncon entries; the lift reward scans them,
   env/sawyer/sawyer_lift_obstacle.py:109-123), as indices into the simulated-geom list */
int orc_dyn_last_contacts(int32_t *g1, int32_t *g2) {
    for (int c = 0; c < g_last_ncon; c++) { g1[c] = g_last_g1[c]; g2[c] = g_last_g2[c]; }
    return g_last_ncon;
}

/* mj_forward's part that the env reads: kinematics + bias at the current state (no integration) */
int orc_dyn_forward(void *h, const double *qpos, const double *qvel, double *bias, double *xpos, double *xquat) {
    dyn_model *m = (dyn_model *)h;
    dyn_data D;
    memset(&D, 0, sizeof(D));
    double q[256], v[256], zero[DMAXD] = {0}, ctrl[DMAXA] = {0};
    memcpy(q, qpos, sizeof(double) * m->nq); memcpy(v, qvel, sizeof(double) * m->nv);
    int ec = m->enable_contacts;
    m->enable_contacts = 0;
    substep(m, q, v, ctrl, zero, &D, NULL, NULL);
    m->enable_contacts = ec;
    if (bias) memcpy(bias, D.bias, sizeof(double) * m->nd);
    if (xpos) for (int i = 0; i < m->nb; i++) for (int k = 0; k < 3; k++) xpos[3 * i + k] = D.xpos[i][k];
    if (xquat) for (int i = 0; i < m->nb; i++) for (int k = 0; k < 4; k++) xquat[4 * i + k] = D.xquat[i][k];
    return 0;
}

/* joint-space inertia, bias force and body COMs at a state (unit tests) */
int orc_dyn_mass_bias(void *h, const double *qpos, const double *qvel, double *M, double *bias, double *com) {
    dyn_model *m = (dyn_model *)h;
    dyn_data *D = (dyn_data *)calloc(1, sizeof(dyn_data));
    double q[256], v[256], zero[DMAXD] = {0}, ctrl[DMAXA] = {0};
    memcpy(q, qpos, sizeof(double) * m->nq); memcpy(v, qvel, sizeof(double) * m->nv);
    int ec = m->enable_contacts;
    m->enable_contacts = 0;
    substep(m, q, v, ctrl, zero, D, NULL, NULL);
    m->enable_contacts = ec;
    memcpy(M, D->M, sizeof(double) * m->nd * m->nd);
    memcpy(bias, D->bias, sizeof(double) * m->nd);
    for (int i = 0; i < m->nb; i++) for (int k = 0; k < 3; k++) com[3 * i + k] = D->com[i][k];
    free(D);
    return 0;
}
