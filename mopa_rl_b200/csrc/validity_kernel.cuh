// State-validity kernel pieces shared by the batch checker and the RRT-Connect kernel.
//
// Hot function replaced: MujocoStateValidityChecker::isValid
// (motion_planners/src/mujoco_ompl_interface.cpp:909-978).  Per query the reference runs a
// full mj_fwdPosition under a mutex; here one thread runs the forward kinematics of just
// the bodies that carry collidable geoms, keeps the chain in registers, and leaves the
// world frames of the moving geoms in shared memory, after which candidate pairs are
// culled by bounding spheres and the survivors go to the narrow phase.
#pragma once
#include "scene.h"

namespace mopa {

struct SceneView {  // pointers into the shared-memory copy of the scene blob
    const SceneHeader *H;
    const FkBody *bodies;
    const FkJoint *joints;
    const FkGeom *geoms;
    const ConstFrame *consts;
    const GeomRec *recs;
    const PairRec *pairs;
    const CullEntry *cull;
    const CullGroup *groups;
    const uint16_t *real;
    const uint4 *wmask;
};

__device__ __forceinline__ SceneView view_scene(const unsigned char *blob) {
    SceneView v;
    v.H = reinterpret_cast<const SceneHeader *>(blob);
    v.bodies = reinterpret_cast<const FkBody *>(blob + v.H->off_body);
    v.joints = reinterpret_cast<const FkJoint *>(blob + v.H->off_joint);
    v.geoms = reinterpret_cast<const FkGeom *>(blob + v.H->off_geom);
    v.consts = reinterpret_cast<const ConstFrame *>(blob + v.H->off_const);
    v.recs = reinterpret_cast<const GeomRec *>(blob + v.H->off_rec);
    v.pairs = reinterpret_cast<const PairRec *>(blob + v.H->off_pair);
    v.cull = reinterpret_cast<const CullEntry *>(blob + v.H->off_cull);
    v.groups = reinterpret_cast<const CullGroup *>(blob + v.H->off_group);
    v.real = reinterpret_cast<const uint16_t *>(blob + v.H->off_real);
    v.wmask = reinterpret_cast<const uint4 *>(blob + v.H->off_wmask);
    return v;
}

// cull test of one pair for the state whose frames sit at frames[slot * stride + lane] (see scene.h: CullKind)
__device__ __forceinline__ bool cull_survives(const PairRec &pr, const CullEntry &e, const float *frames, int stride, int lane) {
    if (pr.ckind == CK_NONE) return true;
    const float *fa = frames + (size_t)pr.anchor_slot * stride + lane;
    const V3 ac{fa[0], fa[stride], fa[2 * stride]};
    if (pr.ckind == CK_PLANE) return !(dot(V3{e.x, e.y, e.z}, ac) > e.w);
    V3 pc{e.x, e.y, e.z};
    if (pr.ckind == CK_SPHERE_MOVING) {
        const float *fp = frames + (size_t)pr.partner_slot * stride + lane;
        pc = V3{fp[0], fp[stride], fp[2 * stride]};
    }
    const V3 d = pc - ac;
    return !(dot(d, d) > e.w);
}

struct Frame { V3 pos; Q4 quat; M3 mat; };
struct Pose { V3 pos; Q4 quat; };

// Forward kinematics of the collision-relevant sub-tree for one state.  `q` is the qpos
// row (fp32).  World frames of moving geoms are written to frames[(slot+k)*stride + lane].
// The running frame lives in registers; a body whose parent is not its predecessor in the (depth-first) body order restarts
// from one of two saved poses or from the constant frame of a world-welded parent.  Those restarts and saves are rare and the
// same for every lane, so they sit behind real branches instead of register selects on every body.
template <class QPos>
__device__ __forceinline__ void fk_state(const SceneView &S, const QPos &q, float *frames, int stride, int lane) {
    Frame cur;
    Pose s0, s1;
    cur.pos = V3{0, 0, 0}; cur.quat = Q4{1, 0, 0, 0}; cur.mat = q2m(cur.quat);
    s0.pos = cur.pos; s0.quat = cur.quat; s1 = s0;
    const int nb = S.H->n_body;
#pragma unroll 1
    for (int b = 0; b < nb; b++) {
        const FkBody B = S.bodies[b];
        if (B.parent_sel != SEL_CUR) {
            if (B.parent_sel == SEL_CONST) {
                const ConstFrame &c = S.consts[B.const_idx];
                cur.pos = V3{c.px, c.py, c.pz};
                cur.quat = Q4{c.qw, c.qx, c.qy, c.qz};
#pragma unroll
                for (int k = 0; k < 9; k++) cur.mat.m[k] = c.m[k];
            } else {
                const Pose &s = B.parent_sel == SEL_SLOT0 ? s0 : s1;
                cur.pos = s.pos; cur.quat = s.quat;
                cur.mat = q2m(cur.quat);   // what the saved body computed
            }
        }
        V3 pos = cur.pos + mulMV(cur.mat, V3{B.px, B.py, B.pz});
        Q4 quat = qmul(cur.quat, Q4{B.qw, B.qx, B.qy, B.qz});
        for (int j = B.jnt_begin; j < B.jnt_end; j++) {
            const FkJoint J = S.joints[j];
            if (J.type == J_HINGE) {
                V3 anchor = pos;
                V3 jp{J.jx, J.jy, J.jz};
                if (J.has_jpos) anchor = pos + mulMV(q2m(quat), jp);
                float sn, cs;
                sincos_cw((q(J.qadr) - J.qpos0) * 0.5f, &sn, &cs);
                quat = qmul(quat, Q4{cs, sn * J.ax, sn * J.ay, sn * J.az});
                pos = anchor;
                if (J.has_jpos) pos = anchor - mulMV(q2m(quat), jp);
            } else if (J.type == J_SLIDE) {
                V3 ax = mulMV(q2m(quat), V3{J.ax, J.ay, J.az});
                float dq = q(J.qadr) - J.qpos0;
                pos = madd(pos, dq, ax);
            } else {  // free
                pos = V3{q(J.qadr), q(J.qadr + 1), q(J.qadr + 2)};
                float w = q(J.qadr + 3), x = q(J.qadr + 4), y = q(J.qadr + 5), z = q(J.qadr + 6);
                float n = sqrtf(fmaf(z, z, fmaf(y, y, fmaf(x, x, w * w))));
                quat = Q4{w / n, x / n, y / n, z / n};
            }
        }
        cur.pos = pos; cur.quat = quat; cur.mat = q2m(quat);
        if (B.save_sel >= 0) {
            if (B.save_sel == 0) { s0.pos = pos; s0.quat = quat; }
            else { s1.pos = pos; s1.quat = quat; }
        }
        for (int g = B.geom_begin; g < B.geom_end; g++) {
            const FkGeom &G = S.geoms[g];
            V3 gp = cur.pos + mulMV(cur.mat, V3{G.px, G.py, G.pz});
            float *f = frames + (size_t)G.slot * stride + lane;
            f[0] = gp.x; f[stride] = gp.y; f[2 * stride] = gp.z;
            if (G.kind >= K_BOX) {   // boxes and mesh hulls need the full frame
                M3 L;
#pragma unroll
                for (int k = 0; k < 9; k++) L.m[k] = G.m[k];
                M3 W = mulMM(cur.mat, L);
#pragma unroll
                for (int k = 0; k < 9; k++) f[(3 + k) * stride] = W.m[k];
            } else if (G.kind != K_SPHERE) {
                V3 a = mulMV(cur.mat, V3{G.m[2], G.m[5], G.m[8]});
                f[3 * stride] = a.x; f[4 * stride] = a.y; f[5 * stride] = a.z;
            }
        }
    }
}

template <bool MESH>
__device__ __forceinline__ void load_geom(Geom &g, const GeomRec &r, const float *frames, int stride, int lane) {
    g.kind = r.kind;
    g.size = V3{r.sx, r.sy, r.sz};
    g.hull = nullptr; g.nhull = 0;
    if (MESH && r.kind == K_MESH) {   // sx / sy carry (byte offset of the hull vertices from this record, vertex count)
        g.hull = reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(&r) + __float_as_int(r.sx));
        g.nhull = __float_as_int(r.sy);
        g.size.z = r.rbound;   // bounding radius of the hull (plane - hull cull)
    }
    if (r.slot < 0) {
        g.c = V3{r.px, r.py, r.pz};
#pragma unroll
        for (int k = 0; k < 9; k++) g.R.m[k] = r.m[k];
        return;
    }
    const float *f = frames + (size_t)r.slot * stride + lane;
    g.c = V3{f[0], f[stride], f[2 * stride]};
#pragma unroll
    for (int k = 0; k < 9; k++) g.R.m[k] = 0.0f;
    if (r.kind >= K_BOX) {
#pragma unroll
        for (int k = 0; k < 9; k++) g.R.m[k] = f[(3 + k) * stride];
    } else if (r.kind != K_SPHERE) {
        g.R.m[2] = f[3 * stride]; g.R.m[5] = f[4 * stride]; g.R.m[8] = f[5 * stride];
    }
}

// signed distance of a deferred (box-box / MPR) pair
template <bool MESH>
__device__ __forceinline__ float heavy_dist(int cls, const Geom &a, const Geom &b) {
    if (cls == PC_BOX_BOX) return box_box(a, b);
    if (mpr_certainly_separate(a, b)) return MOPA_BIG;
    float depth;
    if (mpr_penetration<MESH>(a, b, &depth)) return -depth;
    return MOPA_BIG;
}

}  // namespace mopa
