/* Flat scene description handed across the C ABI.
 *
 * Replaces the `mjModel*` that the reference obtains from mj_loadXML
 * (motion_planners/include/mujoco_wrapper.h:75-89, KinematicPlanner.cpp:62-80).  MuJoCo
 * is not available, so the host compiles the MJCF subset itself
 * (mopa_rl_b200/mjcf.py) and passes the arrays below.  Field names follow mjModel.
 * All floating-point arrays are float64; each implementation converts to its own
 * arithmetic type on load.  Pointers are only read during the *_create call.
 */
#ifndef MOPA_MODEL_DESC_H
#define MOPA_MODEL_DESC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { MOPA_GEOM_PLANE = 0, MOPA_GEOM_SPHERE = 2, MOPA_GEOM_CAPSULE = 3, MOPA_GEOM_CYLINDER = 5, MOPA_GEOM_BOX = 6, MOPA_GEOM_MESH = 7 };
enum { MOPA_JNT_FREE = 0, MOPA_JNT_BALL = 1, MOPA_JNT_SLIDE = 2, MOPA_JNT_HINGE = 3 };

typedef struct mopa_model_desc {
    int32_t nq, nv, nbody, njnt, ngeom, nsite, nexclude, nu;
    /* bodies */
    const int32_t *body_parentid;  /* [nbody] */
    const int32_t *body_weldid;    /* [nbody] */
    const int32_t *body_jntadr;    /* [nbody] first joint or -1 */
    const int32_t *body_jntnum;    /* [nbody] */
    const double  *body_pos;       /* [nbody][3] */
    const double  *body_quat;      /* [nbody][4] wxyz */
    /* joints */
    const int32_t *jnt_type;       /* [njnt] mjtJoint */
    const int32_t *jnt_qposadr;    /* [njnt] */
    const int32_t *jnt_dofadr;     /* [njnt] */
    const int32_t *jnt_bodyid;     /* [njnt] */
    const int32_t *jnt_limited;    /* [njnt] */
    const double  *jnt_pos;        /* [njnt][3] */
    const double  *jnt_axis;       /* [njnt][3] */
    const double  *jnt_range;      /* [njnt][2] */
    const double  *qpos0;          /* [nq] */
    /* geoms */
    const int32_t *geom_type;      /* [ngeom] mjtGeom */
    const int32_t *geom_bodyid;    /* [ngeom] */
    const int32_t *geom_contype;   /* [ngeom] */
    const int32_t *geom_conaffinity;
    const double  *geom_pos;       /* [ngeom][3] */
    const double  *geom_quat;      /* [ngeom][4] */
    const double  *geom_size;      /* [ngeom][3] */
    const double  *geom_margin;    /* [ngeom] */
    const double  *geom_rbound;    /* [ngeom] */
    /* <contact><exclude> body pairs */
    const int32_t *exclude_body;   /* [nexclude][2] */
    /* sites */
    const int32_t *site_bodyid;    /* [nsite] */
    const double  *site_pos;       /* [nsite][3] */
    const double  *site_quat;      /* [nsite][4] */
    /* collision meshes: convex-hull vertices in the geom frame (mjModel mesh_vert of the hull;
       MuJoCo 2.0 collides mesh geoms through their convex hull with libccd).  geom_dataid is the
       mesh of a MOPA_GEOM_MESH geom, -1 otherwise. */
    int32_t nmesh, nmeshvert;
    const int32_t *geom_dataid;    /* [ngeom] */
    const int32_t *mesh_vertadr;   /* [nmesh] */
    const int32_t *mesh_vertnum;   /* [nmesh] */
    const double  *mesh_vert;      /* [nmeshvert][3] */
} mopa_model_desc;

#ifdef __cplusplus
}
#endif
#endif
