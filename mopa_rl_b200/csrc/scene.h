// Device-side scene tables for the state-validity kernel.
//
// Built once per planner (host, scene_build.cu) from the flat mjModel-style arrays; plays
// the role of the mjModel/mjData pair the reference planner owns
// (motion_planners/KinematicPlanner.cpp:62-95).  Everything the kernel needs per query is
// derived from qpos on chip; nothing here depends on the query.
#pragma once
#include <stdint.h>
#include <vector>

#include "collide.cuh"

namespace mopa {

enum { SEL_CUR = 0, SEL_SLOT0 = 1, SEL_SLOT1 = 2, SEL_CONST = 3 };
enum { SAVE_NONE = -1 };
enum { J_FREE = 0, J_BALL = 1, J_SLIDE = 2, J_HINGE = 3 };

struct FkJoint {
    int type, qadr, has_jpos, pad;
    float qpos0, ax, ay, az;
    float jx, jy, jz, pad2;
};
struct FkBody {  // one body of the kinematic sub-tree that carries collidable geoms (DFS order)
    float px, py, pz, pad0;       // body_pos
    float qw, qx, qy, qz;         // body_quat
    int parent_sel, const_idx, save_sel, pad1;
    int jnt_begin, jnt_end, geom_begin, geom_end;
};
struct FkGeom {  // collidable geom attached to an FkBody
    float px, py, pz;
    int kind;
    float m[9];
    int slot;     // float offset of this geom's world frame in the per-query frame store
    int rec;      // index into GeomRec
    int pad;
};
struct ConstFrame {  // world frame of a static (world-welded) parent body
    float px, py, pz, pad;
    float qw, qx, qy, qz;
    float m[9];
    float pad2[3];
};
struct GeomRec {  // every collidable geom that appears in at least one candidate pair
    float sx, sy, sz;   // geom_size; K_MESH: int bits of (byte offset from this record to the hull vertices, count)
    int kind;
    int slot;        // >= 0: moving geom, float offset in the frame store;  -1: static
    float px, py, pz;  // world frame when static
    float m[9];
    float rbound;
    int geom_id;     // mjModel geom id
    int pad;
};
// Candidate pairs, in kernel order: sorted by the kind of cull test, then by anchor (the first moving geom).  Two parallel arrays,
// 16 bytes per pair each: what the narrow phase needs (PairRec) and what the cull sweep needs (CullEntry).  Runs of pairs with the
// same (cull kind, anchor) are described by CullGroup records: the sweep loads the anchor centre once per run and has branch-free
// inner loops.  The arrays are laid out in windows of 32 entries (one survivor mask word per window): a run never straddles a
// window (longer runs are split), runs are padded to multiples of four and windows are filled up with dummy pairs (cls = PC_NONE)
// that never survive a cull test.  `real` lists the indices of the non-dummy entries for consumers that walk the pairs one by one.
// Pairs of a moving and a static geom whose bounding spheres can never touch, whatever the joint angles, are dropped when the
// scene is built (scene_build.cu: reach spheres).
struct PairRec {
    uint16_t anchor_slot, partner_slot;  // frame-store offsets; partner_slot == 0xFFFF: static partner
    uint16_t ga, gb;    // GeomRec indices, kind(ga) <= kind(gb)  (the order the narrowphase expects)
    uint16_t canon;     // index in the canonical (g1<g2 lexicographic) candidate list
    uint8_t cls;        // PairClass
    uint8_t ckind;      // CullKind
    uint8_t mkey;       // portal-refinement pairs: rank (< 16) of the pair's (kind, kind) combination among those of the scene;
    uint8_t pad[3];     //   the queued refinements are grouped by it so that the lanes of a warp run the same support routines
};
enum CullKind : int {
    CK_SPHERE_STATIC = 0,   // e = (partner centre, squared cull distance): cull when |centre - anchor centre|^2 > e.w
    CK_SPHERE_MOVING = 1,   // e.x = int bits of the partner's frame-store offset, e.w = squared cull distance
    CK_PLANE = 2,           // e = (unit normal of the static plane, offset): cull when dot(normal, anchor centre) > e.w
    CK_NONE = 3             // always passed to the narrow phase
};
struct alignas(16) CullEntry { float x, y, z, w; };
struct CullGroup {
    uint16_t anchor_slot;   // frame-store offset of the anchor centre
    uint8_t kind, count;    // CullKind; entries in the run (multiple of four, <= 32)
    uint8_t bitpos, flush;  // position of the run's first entry in its window; last run of the window
    uint16_t first;         // index of the run's first entry
};

struct SceneHeader {
    int nq, nq4;             // qpos row length, and in float4 units (row stride = 4*nq4 floats)
    int n_body, n_joint, n_geom, n_const, n_rec, n_pair;
    int frame_floats;        // floats per query in the frame store
    int n_site_pad;
    float threshold;
    int off_body, off_joint, off_geom, off_const, off_rec, off_pair;  // byte offsets from blob start
    int blob_bytes;
    int off_hull;            // hull vertices of collision meshes (float xyz triplets)
    int n_hull_vert;
    int off_cull, off_group, n_group;   // CullEntry[n_pair] (parallel to the pair records), CullGroup[n_group]
    int off_real, n_real;               // uint16 indices of the non-dummy pair entries
    int off_wmask;                      // uint32 [n_pair / 32][4]: per window, which entries belong to work list 0 / 1 / 2 (analytic, box-box, portal refinement)
    int n_pruned;                       // candidate pairs dropped at build time (bounding spheres out of reach for every joint configuration)
};

struct HostScene {
    SceneHeader hdr;
    std::vector<unsigned char> blob;   // header + tables, 16-byte aligned sections
    std::vector<int> canon_g1, canon_g2;  // canonical pair list (mjModel geom ids)
};

struct mopa_model_desc_fwd;

}  // namespace mopa
