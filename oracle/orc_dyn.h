/* TEST INFRASTRUCTURE — shared types of the physics-step oracle (orc_dyn.c, orc_contact.c). */
#ifndef ORC_DYN_H
#define ORC_DYN_H
#include "../include/mopa_dyn_desc.h"

#define DMAXB 24
#define DMAXD 24
#define DMAXA 16
#define DMAXG 64
#define DMAXC 36 /* constraint rows (same cap as the kernel) */
#define MINVAL 1e-15

typedef struct { double w[3], v[3]; } sv6; /* spatial motion (w, v_O) or force (n_O, f) about the world origin */
typedef struct { double m, h[3], I[9]; } sinert; /* mass, m*c, inertia about the origin */

typedef struct {
    int nq, nv, nb, nd, nact, ngeom, npair, iterations;
    double h, g[3], tolerance;
    int b_parent[DMAXB], b_jtype[DMAXB], b_qadr[DMAXB], b_vadr[DMAXB], b_dadr[DMAXB];
    double b_pos[DMAXB][3], b_quat[DMAXB][4], b_rootpos[DMAXB][3], b_rootquat[DMAXB][4], b_jaxis[DMAXB][3], b_jpos[DMAXB][3],
        b_qpos0[DMAXB], b_mass[DMAXB], b_ipos[DMAXB][3], b_iquat[DMAXB][4], b_inertia[DMAXB][3];
    int d_body[DMAXD], d_qadr[DMAXD], d_vadr[DMAXD], d_limited[DMAXD], d_parent[DMAXD];
    double d_armature[DMAXD], d_damping[DMAXD], d_range[DMAXD][2], d_solref[DMAXD][2], d_solimp[DMAXD][5], d_margin[DMAXD];
    int a_dof[DMAXA], a_kind[DMAXA], a_ctrllimited[DMAXA], a_forcelimited[DMAXA];
    double a_kp[DMAXA], a_kv[DMAXA], a_gear[DMAXA], a_ctrlrange[DMAXA][2], a_forcerange[DMAXA][2];
    int enable_contacts, max_rows, integrator; /* integrator: 0 semi-implicit Euler, 1 RK4 (oracle-side switch, orc_dyn_set_integrator) */
    /* contact geoms (orc_contact.c) */
    int g_body[DMAXG], g_type[DMAXG], g_condim[DMAXG];
    double g_pos[DMAXG][3], g_quat[DMAXG][4], g_size[DMAXG][3], g_rbound[DMAXG], g_margin[DMAXG], g_friction[DMAXG][3],
        g_solref[DMAXG][2], g_solimp[DMAXG][5];
    int *p_g1, *p_g2;
} dyn_model;

/* per-step scratch that the env layer reads back (kinematics at the start of the last substep,
   which is what mjData holds after mj_step) */
typedef struct {
    double xpos[DMAXB][3], xquat[DMAXB][4], xmat[DMAXB][9];
    double bias[DMAXD];
    int ncon;
    int con_pair[DMAXC / 3 + 1]; /* candidate-pair index of every contact of the last mj_step */
    double cforce; /* sum over contacts of |f_n| + |f_t1| + |f_t2| of the last mj_step (BaseEnv.get_contact_force) */
    double M[DMAXD * DMAXD], com[DMAXB][3]; /* debug / unit tests */
} dyn_data;


typedef struct {
    double J[DMAXD];
    double pos, margin, solref[2], solimp[5];
    int type; /* 0 limit, 1 contact normal, 2 tangent */
    double mu;
    int sig;  /* identity of the row across substeps (warm start): limits -(2*dof+side+1), contacts pair*16 + point*4 + dir */
} crow;
typedef struct { int have_a; double a[DMAXD]; } warm_t; /* qacc of the previous substep (mjData.qacc_warmstart) */
int orc_contact_rows(const dyn_model *m, const dyn_data *d, const sv6 *S, crow *rows, int maxrows);
#endif
