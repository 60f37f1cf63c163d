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
constexpr int DMAXGM = 8;  // moving contact geoms with a non-identity local rotation
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
    int p_g1[640], p_g2[640];
    double g_mat[DMAXG][9];   // rotation matrix: world (static geom) or local (moving geom)
    int g_mslot[DMAXG];       // -2 static, -1 moving with identity local rotation, >= 0 slot in the per-substep world-matrix cache
    int gm_geom[DMAXGM], ngm;
    int integrator;           // 0 Euler, 1 RK4
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

}  // namespace mopa
