/* Flat description of the simulated part of a scene (what mj_step integrates).
 *
 * The reference advances a full mjModel with MuJoCo's mj_step (env/base.py:388-400,
 * env/sawyer/sawyer_push_obstacle.py:186-203).  Here the host (mopa_rl_b200/dynmodel.py) selects
 * the kinematic trees that can move under forces - trees that contain an actuated joint or a
 * collidable geom - and passes them as flat float64 arrays.  The "ghost" indicator/target arms
 * (contype = conaffinity = 0, no actuators) and the target slider are not simulated: nothing
 * can act on them, so their qpos stays where reset put it.
 */
#ifndef MOPA_DYN_DESC_H
#define MOPA_DYN_DESC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mopa_dyn_desc {
    int32_t nq, nv;           /* sizes of the FULL qpos / qvel rows the state arrays use */
    int32_t nb, nd, nact;     /* simulated bodies, dofs, actuators */
    int32_t ngeom, npair;     /* contact geoms and candidate contact pairs */
    int32_t iterations;       /* constraint solver sweeps (option iterations) */
    double timestep;
    double gravity[3];
    double tolerance;         /* solver early-termination threshold (option tolerance) */
    /* simulated bodies, parents first */
    const int32_t *b_parent;  /* [nb] index of the simulated parent, -1: static parent */
    const int32_t *b_bodyid;  /* [nb] mjModel body id */
    const double *b_pos;      /* [nb][3] body_pos */
    const double *b_quat;     /* [nb][4] body_quat */
    const double *b_rootpos;  /* [nb][3] world frame of the static parent (used when b_parent < 0) */
    const double *b_rootquat; /* [nb][4] */
    const int32_t *b_jtype;   /* [nb] mjtJoint of the body's joint, -1: welded to its parent */
    const int32_t *b_qadr;    /* [nb] qpos address of the joint */
    const int32_t *b_vadr;    /* [nb] qvel address of the joint */
    const int32_t *b_dadr;    /* [nb] first simulated-dof index of the joint */
    const double *b_jaxis;    /* [nb][3] */
    const double *b_jpos;     /* [nb][3] */
    const double *b_qpos0;    /* [nb] joint reference */
    const double *b_mass;     /* [nb] */
    const double *b_ipos;     /* [nb][3] */
    const double *b_iquat;    /* [nb][4] */
    const double *b_inertia;  /* [nb][3] */
    /* simulated dofs */
    const int32_t *d_body;    /* [nd] simulated-body index */
    const int32_t *d_qadr;    /* [nd] qpos address (-1 for the rotational dofs of a free joint) */
    const int32_t *d_vadr;    /* [nd] qvel address */
    const double *d_armature; /* [nd] */
    const double *d_damping;  /* [nd] */
    const int32_t *d_limited; /* [nd] */
    const double *d_range;    /* [nd][2] */
    const double *d_solref;   /* [nd][2] limit solref */
    const double *d_solimp;   /* [nd][5] limit solimp */
    const double *d_margin;   /* [nd] */
    /* actuators (joint transmission) */
    const int32_t *a_dof;     /* [nact] simulated-dof index */
    const int32_t *a_kind;    /* [nact] 0 motor, 1 position, 2 velocity */
    const int32_t *a_ctrllimited;
    const int32_t *a_forcelimited;
    const double *a_kp, *a_kv, *a_gear;
    const double *a_ctrlrange;  /* [nact][2] */
    const double *a_forcerange; /* [nact][2] */
    /* contact geoms */
    const int32_t *g_body;    /* [ngeom] simulated-body index, -1: static */
    const int32_t *g_geomid;  /* [ngeom] mjModel geom id */
    const int32_t *g_type;    /* [ngeom] mjtGeom */
    const double *g_pos;      /* [ngeom][3] local frame (static geoms: world frame) */
    const double *g_quat;     /* [ngeom][4] */
    const double *g_size;     /* [ngeom][3] */
    const double *g_rbound;   /* [ngeom] */
    const double *g_margin;   /* [ngeom] */
    const double *g_friction; /* [ngeom][3] */
    const double *g_solref;   /* [ngeom][2] */
    const double *g_solimp;   /* [ngeom][5] */
    const int32_t *g_condim;  /* [ngeom] */
    const int32_t *p_g1;      /* [npair] indices into the contact geoms */
    const int32_t *p_g2;
    int32_t integrator;       /* <option integrator>: 0 Euler (semi-implicit, implicit joint damping), 1 RK4 (mj_RungeKutta, N = 4) */
    int32_t pad_;
} mopa_dyn_desc;

#ifdef __cplusplus
}
#endif
#endif
