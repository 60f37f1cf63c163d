// Host-side scene compiler: mjModel-style arrays -> device tables for the validity kernel.
//
// Covers what the reference does at planner construction
// (KinematicPlanner.cpp:42-120: load model, build the state space over active joints,
// install MujocoStateValidityChecker with `ignored_contacts` and `contact_threshold`) plus
// the query-independent half of MuJoCo's collision pipeline: the
// candidate-pair filters and the world frames of world-welded geoms.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/mopa_model_desc.h"
#include "scene.h"

namespace mopa {

struct HFrame { V3 pos; Q4 quat; M3 mat; };

static int kind_of(int mjtype) {
    switch (mjtype) {
    case MOPA_GEOM_PLANE: return K_PLANE;
    case MOPA_GEOM_SPHERE: return K_SPHERE;
    case MOPA_GEOM_CAPSULE: return K_CAPSULE;
    case MOPA_GEOM_CYLINDER: return K_CYLINDER;
    case MOPA_GEOM_BOX: return K_BOX;
    case MOPA_GEOM_MESH: return K_MESH;
    default: return -1;
    }
}

static bool pair_allowed(const mopa_model_desc *d, int g1, int g2) {
    int b1 = d->geom_bodyid[g1], b2 = d->geom_bodyid[g2];
    int w1 = d->body_weldid[b1], w2 = d->body_weldid[b2];
    if (w1 == w2) return false;
    if (w1 != 0 && w2 != 0) {
        int wp1 = d->body_weldid[d->body_parentid[w1]], wp2 = d->body_weldid[d->body_parentid[w2]];
        if (wp1 == w2 || wp2 == w1) return false;
    }
    for (int e = 0; e < d->nexclude; e++) {
        int e1 = d->exclude_body[2 * e], e2 = d->exclude_body[2 * e + 1];
        if ((e1 == b1 && e2 == b2) || (e1 == b2 && e2 == b1)) return false;
    }
    bool mask = (d->geom_contype[g1] & d->geom_conaffinity[g2]) || (d->geom_contype[g2] & d->geom_conaffinity[g1]);
    if (!mask) return false;
    if (d->geom_type[g1] == MOPA_GEOM_PLANE && d->geom_type[g2] == MOPA_GEOM_PLANE) return false;
    return true;
}

template <class T>
static int append(std::vector<unsigned char> &blob, const std::vector<T> &v) {
    while (blob.size() % 16) blob.push_back(0);
    int off = (int)blob.size();
    const unsigned char *p = (const unsigned char *)v.data();
    blob.insert(blob.end(), p, p + sizeof(T) * v.size());
    return off;
}

// ---- reach spheres.  Where can the centre of a geom be, whatever the joint values?  Walking from the geom's body up to `stop`
// (exclusive), a sphere (c, r) in the current body's frame is carried through the body's joints - a hinge sweeps the centre on a
// circle around its axis (any angle), a limited slide moves it along its axis (three times the joint range is allowed for,
// soft limits are overshot in simulation), free / ball / unlimited slide joints give up - and through the body's fixed offset
// into the parent's frame.  The result bounds the centre in the frame of `stop` (after stop's own joints).  Double precision;
// the callers add a millimetre of slack.
struct Reach { double c[3]; double r; bool bounded; };
static void quat_rot(const double *q, const double *v, double *out) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double t[3] = {2 * (y * v[2] - z * v[1]), 2 * (z * v[0] - x * v[2]), 2 * (x * v[1] - y * v[0])};
    out[0] = v[0] + w * t[0] + (y * t[2] - z * t[1]);
    out[1] = v[1] + w * t[1] + (z * t[0] - x * t[2]);
    out[2] = v[2] + w * t[2] + (x * t[1] - y * t[0]);
}
static Reach reach_of_geom(const mopa_model_desc *d, int g, int stop) {
    Reach R;
    R.bounded = true; R.r = 0.0;
    for (int k = 0; k < 3; k++) R.c[k] = d->geom_pos[3 * g + k];
    for (int b = d->geom_bodyid[g]; b != stop; b = d->body_parentid[b]) {
        if (b == 0) { R.bounded = false; return R; }   // stop is not an ancestor
        for (int k = d->body_jntnum[b] - 1; k >= 0; k--) {
            const int j = d->body_jntadr[b] + k;
            double ax[3] = {d->jnt_axis[3 * j], d->jnt_axis[3 * j + 1], d->jnt_axis[3 * j + 2]};
            const double an = sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
            if (d->jnt_type[j] == MOPA_JNT_HINGE && an > 0) {
                double rel[3], along = 0;
                for (int c = 0; c < 3; c++) { ax[c] /= an; rel[c] = R.c[c] - d->jnt_pos[3 * j + c]; along += rel[c] * ax[c]; }
                double rho2 = 0;
                for (int c = 0; c < 3; c++) { const double perp = rel[c] - along * ax[c]; rho2 += perp * perp; R.c[c] = d->jnt_pos[3 * j + c] + along * ax[c]; }
                R.r += sqrt(rho2);
            } else if (d->jnt_type[j] == MOPA_JNT_SLIDE && an > 0 && d->jnt_limited[j]) {
                const double lo = d->jnt_range[2 * j], hi = d->jnt_range[2 * j + 1], mid = 0.5 * (lo + hi) - d->qpos0[d->jnt_qposadr[j]];
                for (int c = 0; c < 3; c++) R.c[c] += ax[c] / an * mid;
                R.r += 1.5 * (hi - lo);
            } else { R.bounded = false; return R; }
        }
        double w[3];
        quat_rot(d->body_quat + 4 * b, R.c, w);
        for (int c = 0; c < 3; c++) R.c[c] = d->body_pos[3 * b + c] + w[c];
    }
    return R;
}
static int common_ancestor(const mopa_model_desc *d, int b1, int b2) {
    for (int a = b1;; a = d->body_parentid[a]) {
        for (int b = b2;; b = d->body_parentid[b]) {
            if (a == b) return a;
            if (b == 0) break;
        }
        if (a == 0) return 0;
    }
}
// true when the bounding spheres of the two geoms (radius sum + margin, MuJoCo's filter) cannot touch for any joint configuration
static bool never_in_reach(const mopa_model_desc *d, int g1, int g2, double margin) {
    const int stop = common_ancestor(d, d->geom_bodyid[g1], d->geom_bodyid[g2]);
    const Reach A = reach_of_geom(d, g1, stop), B = reach_of_geom(d, g2, stop);
    if (!A.bounded || !B.bounded) return false;
    double dist2 = 0;
    for (int c = 0; c < 3; c++) dist2 += (A.c[c] - B.c[c]) * (A.c[c] - B.c[c]);
    return sqrt(dist2) - A.r - B.r > d->geom_rbound[g1] + d->geom_rbound[g2] + margin + 1e-3;
}

void build_scene(const mopa_model_desc *d, const int32_t *ignored, int nignored, double threshold, HostScene &out) {
    const int nb = d->nbody, ng = d->ngeom;
    // ---- canonical candidate list
    out.canon_g1.clear();
    out.canon_g2.clear();
    for (int g1 = 0; g1 < ng; g1++)
        for (int g2 = g1 + 1; g2 < ng; g2++) {
            if (!pair_allowed(d, g1, g2)) continue;
            bool ign = false;
            for (int k = 0; k < nignored; k++)
                if (ignored[2 * k] == g1 && ignored[2 * k + 1] == g2) ign = true;
            if (ign) continue;
            if (kind_of(d->geom_type[g1]) < 0 || kind_of(d->geom_type[g2]) < 0)
                throw std::runtime_error("collidable geom of unsupported type (ellipsoid/hfield) in a candidate pair");
            for (int g : {g1, g2})
                if (d->geom_type[g] == MOPA_GEOM_MESH && (d->geom_dataid[g] < 0 || d->geom_dataid[g] >= d->nmesh))
                    throw std::runtime_error("collidable mesh geom without convex-hull vertices (recompile the scene)");
            out.canon_g1.push_back(g1);
            out.canon_g2.push_back(g2);
        }
    const int npair = (int)out.canon_g1.size();
    if (npair >= 65535) throw std::runtime_error("too many candidate pairs");
    std::vector<char> used(ng, 0);
    for (int p = 0; p < npair; p++) used[out.canon_g1[p]] = used[out.canon_g2[p]] = 1;

    // ---- frames of world-welded bodies (same fp32 operation order as the per-query FK)
    std::vector<HFrame> sframe(nb);
    sframe[0].pos = V3{0, 0, 0};
    sframe[0].quat = Q4{1, 0, 0, 0};
    sframe[0].mat = q2m(sframe[0].quat);
    for (int b = 1; b < nb; b++) {
        if (d->body_weldid[b] != 0) continue;
        const HFrame &P = sframe[d->body_parentid[b]];
        V3 bp{(float)d->body_pos[3 * b], (float)d->body_pos[3 * b + 1], (float)d->body_pos[3 * b + 2]};
        Q4 bq{(float)d->body_quat[4 * b], (float)d->body_quat[4 * b + 1], (float)d->body_quat[4 * b + 2], (float)d->body_quat[4 * b + 3]};
        sframe[b].pos = P.pos + mulMV(P.mat, bp);
        sframe[b].quat = qmul(P.quat, bq);
        sframe[b].mat = q2m(sframe[b].quat);
    }

    // ---- moving bodies that matter: bodies of used geoms and their moving ancestors
    std::vector<char> rel(nb, 0);
    for (int g = 0; g < ng; g++)
        if (used[g]) {
            int b = d->geom_bodyid[g];
            while (b != 0 && d->body_weldid[b] != 0 && !rel[b]) { rel[b] = 1; b = d->body_parentid[b]; }
        }
    std::vector<int> order;
    for (int b = 1; b < nb; b++)
        if (rel[b]) order.push_back(b);

    std::vector<FkBody> bodies;
    std::vector<FkJoint> joints;
    std::vector<FkGeom> geoms;
    std::vector<ConstFrame> consts;
    std::vector<GeomRec> recs;
    std::vector<int> rec_of_geom(ng, -1);
    std::vector<int> const_of_body(nb, -1);
    std::vector<int> slot_of_body(nb, -1);  // register slot currently holding body's frame
    int slot_owner[2] = {-1, -1};
    int frame_floats = 0;

    // last index in `order` at which each body is needed as a parent
    std::vector<int> last_use(nb, -1);
    for (size_t i = 0; i < order.size(); i++) last_use[d->body_parentid[order[i]]] = (int)i;

    for (size_t i = 0; i < order.size(); i++) {
        int b = order[i], p = d->body_parentid[b];
        FkBody B;
        memset(&B, 0, sizeof(B));
        B.px = (float)d->body_pos[3 * b]; B.py = (float)d->body_pos[3 * b + 1]; B.pz = (float)d->body_pos[3 * b + 2];
        B.qw = (float)d->body_quat[4 * b]; B.qx = (float)d->body_quat[4 * b + 1]; B.qy = (float)d->body_quat[4 * b + 2]; B.qz = (float)d->body_quat[4 * b + 3];
        B.const_idx = -1;
        if (d->body_weldid[p] == 0) {  // static parent (incl. world)
            if (const_of_body[p] < 0) {
                ConstFrame c;
                memset(&c, 0, sizeof(c));
                c.px = sframe[p].pos.x; c.py = sframe[p].pos.y; c.pz = sframe[p].pos.z;
                c.qw = sframe[p].quat.w; c.qx = sframe[p].quat.x; c.qy = sframe[p].quat.y; c.qz = sframe[p].quat.z;
                memcpy(c.m, sframe[p].mat.m, sizeof(c.m));
                const_of_body[p] = (int)consts.size();
                consts.push_back(c);
            }
            B.parent_sel = SEL_CONST;
            B.const_idx = const_of_body[p];
        } else if (i > 0 && order[i - 1] == p) {
            B.parent_sel = SEL_CUR;
        } else if (slot_of_body[p] >= 0) {
            B.parent_sel = SEL_SLOT0 + slot_of_body[p];
        } else {
            throw std::runtime_error("kinematic tree needs more than two saved frames");
        }
        // free slots whose owner is no longer needed
        for (int s = 0; s < 2; s++)
            if (slot_owner[s] >= 0 && last_use[slot_owner[s]] <= (int)i) { slot_of_body[slot_owner[s]] = -1; slot_owner[s] = -1; }
        // does b need to be saved?  (needed as a parent by a body that is not the next one)
        B.save_sel = SAVE_NONE;
        bool needed_later = false;
        for (size_t k = i + 2; k < order.size(); k++)
            if (d->body_parentid[order[k]] == b) needed_later = true;
        if (needed_later) {
            int s = slot_owner[0] < 0 ? 0 : (slot_owner[1] < 0 ? 1 : -1);
            if (s < 0) throw std::runtime_error("kinematic tree needs more than two saved frames");
            slot_owner[s] = b;
            slot_of_body[b] = s;
            B.save_sel = s;
        }
        B.jnt_begin = (int)joints.size();
        for (int k = 0; k < d->body_jntnum[b]; k++) {
            int j = d->body_jntadr[b] + k;
            FkJoint J;
            memset(&J, 0, sizeof(J));
            J.type = d->jnt_type[j];
            if (J.type == MOPA_JNT_BALL) throw std::runtime_error("ball joints are not supported");
            J.qadr = d->jnt_qposadr[j];
            J.qpos0 = (float)d->qpos0[J.qadr];
            J.ax = (float)d->jnt_axis[3 * j]; J.ay = (float)d->jnt_axis[3 * j + 1]; J.az = (float)d->jnt_axis[3 * j + 2];
            J.jx = (float)d->jnt_pos[3 * j]; J.jy = (float)d->jnt_pos[3 * j + 1]; J.jz = (float)d->jnt_pos[3 * j + 2];
            J.has_jpos = (J.jx != 0.0f || J.jy != 0.0f || J.jz != 0.0f);
            joints.push_back(J);
        }
        B.jnt_end = (int)joints.size();
        B.geom_begin = (int)geoms.size();
        for (int g = 0; g < ng; g++) {
            if (!used[g] || d->geom_bodyid[g] != b) continue;
            FkGeom G;
            memset(&G, 0, sizeof(G));
            G.px = (float)d->geom_pos[3 * g]; G.py = (float)d->geom_pos[3 * g + 1]; G.pz = (float)d->geom_pos[3 * g + 2];
            Q4 gq{(float)d->geom_quat[4 * g], (float)d->geom_quat[4 * g + 1], (float)d->geom_quat[4 * g + 2], (float)d->geom_quat[4 * g + 3]};
            M3 gm = q2m(gq);
            memcpy(G.m, gm.m, sizeof(G.m));
            G.kind = kind_of(d->geom_type[g]);
            G.slot = frame_floats;
            frame_floats += (G.kind == K_SPHERE) ? 3 : (G.kind >= K_BOX ? 12 : 6);
            G.rec = (int)recs.size();
            GeomRec Rr;
            memset(&Rr, 0, sizeof(Rr));
            Rr.sx = (float)d->geom_size[3 * g]; Rr.sy = (float)d->geom_size[3 * g + 1]; Rr.sz = (float)d->geom_size[3 * g + 2];
            Rr.kind = G.kind; Rr.slot = G.slot; Rr.rbound = (float)d->geom_rbound[g]; Rr.geom_id = g;
            rec_of_geom[g] = G.rec;
            recs.push_back(Rr);
            geoms.push_back(G);
        }
        B.geom_end = (int)geoms.size();
        bodies.push_back(B);
    }
    if (frame_floats >= 65535) throw std::runtime_error("frame store too large");
    // static geoms
    for (int g = 0; g < ng; g++) {
        if (!used[g] || rec_of_geom[g] >= 0) continue;
        int b = d->geom_bodyid[g];
        GeomRec Rr;
        memset(&Rr, 0, sizeof(Rr));
        Rr.sx = (float)d->geom_size[3 * g]; Rr.sy = (float)d->geom_size[3 * g + 1]; Rr.sz = (float)d->geom_size[3 * g + 2];
        Rr.kind = kind_of(d->geom_type[g]); Rr.slot = -1; Rr.rbound = (float)d->geom_rbound[g]; Rr.geom_id = g;
        V3 gp{(float)d->geom_pos[3 * g], (float)d->geom_pos[3 * g + 1], (float)d->geom_pos[3 * g + 2]};
        Q4 gq{(float)d->geom_quat[4 * g], (float)d->geom_quat[4 * g + 1], (float)d->geom_quat[4 * g + 2], (float)d->geom_quat[4 * g + 3]};
        V3 wp = sframe[b].pos + mulMV(sframe[b].mat, gp);
        M3 wm = mulMM(sframe[b].mat, q2m(gq));
        Rr.px = wp.x; Rr.py = wp.y; Rr.pz = wp.z;
        memcpy(Rr.m, wm.m, sizeof(Rr.m));
        rec_of_geom[g] = (int)recs.size();
        recs.push_back(Rr);
    }
    // ---- pair records + cull entries
    struct PairBuild { PairRec P; CullEntry E; int list; };
    std::vector<PairBuild> pb;
    std::vector<int> mpr_combos;
    int n_pruned = 0;
    for (int p = 0; p < npair; p++) {
        int g1 = out.canon_g1[p], g2 = out.canon_g2[p];
        if (d->geom_type[g1] > d->geom_type[g2]) std::swap(g1, g2);  // lower mjtGeom first, ties keep g1<g2
        const GeomRec &A = recs[rec_of_geom[g1]], &B = recs[rec_of_geom[g2]];
        PairBuild X;
        memset(&X, 0, sizeof(X));
        PairRec &P = X.P;
        CullEntry &E = X.E;
        P.ga = (uint16_t)rec_of_geom[g1];
        P.gb = (uint16_t)rec_of_geom[g2];
        P.canon = (uint16_t)p;
        P.cls = (uint8_t)pair_class(A.kind, B.kind);
        if (P.cls == PC_MPR) {
            const int combo = A.kind * 8 + B.kind;
            size_t k = 0;
            while (k < mpr_combos.size() && mpr_combos[k] != combo) k++;
            if (k == mpr_combos.size()) mpr_combos.push_back(combo);
            P.mkey = (uint8_t)(k < 15 ? k : 15);
        }
        X.list = P.cls < PC_BOX_BOX ? 0 : (P.cls == PC_BOX_BOX ? 1 : (P.cls == PC_MPR ? 2 : 3));
        float margin = fmaxf((float)d->geom_margin[g1], (float)d->geom_margin[g2]);
        const GeomRec *anchor = &A, *partner = &B;
        if (A.slot < 0 || (B.slot >= 0 && B.geom_id < A.geom_id)) { anchor = &B; partner = &A; }
        if (anchor->slot < 0) throw std::runtime_error("candidate pair without a moving geom");
        P.anchor_slot = (uint16_t)anchor->slot;
        P.partner_slot = partner->slot < 0 ? 0xFFFF : (uint16_t)partner->slot;
        if (A.kind == K_PLANE) {
            P.ckind = CK_NONE;
            if (A.slot < 0 && anchor == &B) {
                // static plane against a moving geom: nothing of the geom is closer to the plane than its centre height minus
                // its bounding radius.  The kernels skip the pair when that lower bound is clearly positive (1e-4 of slack
                // against rounding; the validity predicate needs dist <= contact_threshold).  Conservative: never changes a result.
                const V3 nrm{A.m[2], A.m[5], A.m[8]};
                E.x = nrm.x; E.y = nrm.y; E.z = nrm.z;
                E.w = dot(nrm, V3{A.px, A.py, A.pz}) + B.rbound + 1e-4f + fmaxf((float)threshold, 0.0f);
                P.ckind = CK_PLANE;
            }
        } else {   // bounding spheres + margin, exactly MuJoCo's (and the oracle's) filter
            float bound = (A.rbound + B.rbound) + margin;
            E.w = bound * bound;
            if (partner->slot < 0) { P.ckind = CK_SPHERE_STATIC; E.x = partner->px; E.y = partner->py; E.z = partner->pz; }
            else { P.ckind = CK_SPHERE_MOVING; const int off = partner->slot; memcpy(&E.x, &off, 4); }
        }
        if (X.list == 3) continue;   // pairs of classes the narrow phase does not know never collide (PC_NONE)
        // out of reach for every joint configuration: the pair can never pass MuJoCo's bounding-sphere filter (plane pairs: the
        // geom's lowest possible point stays above the plane), so it is not a candidate at all
        bool pruned = false;
        if (A.kind == K_PLANE) {
            if (P.ckind == CK_PLANE) {
                const Reach Rb = reach_of_geom(d, g2, 0);
                if (Rb.bounded) {
                    const double h = A.m[2] * (Rb.c[0] - A.px) + A.m[5] * (Rb.c[1] - A.py) + A.m[8] * (Rb.c[2] - A.pz) - Rb.r - B.rbound;
                    pruned = h > 1e-3 + fmax(threshold, 0.0);
                }
            }
        } else
            pruned = never_in_reach(d, g1, g2, margin);
        if (pruned) { n_pruned++; continue; }
        pb.push_back(X);
    }
    // kernel order: runs of (cull kind, anchor) in windows of 32 entries (see scene.h)
    std::stable_sort(pb.begin(), pb.end(), [](const PairBuild &a, const PairBuild &b) {
        if (a.P.ckind != b.P.ckind) return a.P.ckind < b.P.ckind;
        if (a.P.anchor_slot != b.P.anchor_slot) return a.P.anchor_slot < b.P.anchor_slot;
        return a.P.canon < b.P.canon;
    });
    std::vector<PairRec> pairs;
    std::vector<CullEntry> cull;
    std::vector<CullGroup> groups;
    std::vector<uint16_t> real;
    auto push_dummy = [&](int kind, int anchor_slot) {
        PairRec D;
        memset(&D, 0, sizeof(D));
        D.cls = PC_NONE; D.ckind = (uint8_t)(kind == CK_NONE ? CK_PLANE : kind); D.anchor_slot = (uint16_t)anchor_slot; D.partner_slot = (uint16_t)anchor_slot;
        D.canon = 0xFFFF;
        CullEntry E{0.0f, 0.0f, 0.0f, -1.0f};   // |d|^2 > -1 and 0 > -1: culled by every kind of test
        if (D.ckind == CK_SPHERE_MOVING) { const int off = anchor_slot; memcpy(&E.x, &off, 4); }
        pairs.push_back(D); cull.push_back(E);
    };
    for (size_t i = 0; i < pb.size();) {
        size_t j = i;
        while (j < pb.size() && pb[j].P.ckind == pb[i].P.ckind && pb[j].P.anchor_slot == pb[i].P.anchor_slot) j++;
        const int kind = pb[i].P.ckind, anchor = pb[i].P.anchor_slot;
        while (i < j) {   // the run, cut where it would straddle a window (every piece but the last fills its window exactly)
            while (pairs.size() % 4) push_dummy(CK_PLANE, 0);   // after an unpadded CK_NONE run
            const int bitpos = (int)(pairs.size() % 32);
            const int n = (int)std::min<size_t>(j - i, 32 - bitpos), padded = kind == CK_NONE ? n : (n + 3) & ~3;
            CullGroup G;
            memset(&G, 0, sizeof(G));
            G.anchor_slot = (uint16_t)anchor; G.kind = (uint8_t)kind; G.count = (uint8_t)padded; G.bitpos = (uint8_t)bitpos; G.first = (uint16_t)pairs.size();
            for (int k = 0; k < n; k++) { real.push_back((uint16_t)pairs.size()); pairs.push_back(pb[i + k].P); cull.push_back(pb[i + k].E); }
            for (int k = n; k < padded; k++) push_dummy(kind, anchor);
            groups.push_back(G);
            i += n;
        }
    }
    while (pairs.size() % 32) push_dummy(CK_PLANE, 0);
    for (size_t g = 0; g < groups.size(); g++)   // the last run of every window flushes the survivor mask
        groups[g].flush = (g + 1 == groups.size() || groups[g + 1].bitpos == 0) ? 1 : 0;
    if (pairs.size() >= 65535) throw std::runtime_error("too many candidate pairs");

    // ---- pack
    SceneHeader H;
    memset(&H, 0, sizeof(H));
    H.nq = d->nq;
    H.nq4 = (d->nq + 3) / 4;
    H.n_body = (int)bodies.size(); H.n_joint = (int)joints.size(); H.n_geom = (int)geoms.size();
    H.n_const = (int)consts.size(); H.n_rec = (int)recs.size(); H.n_pair = (int)pairs.size();
    H.frame_floats = frame_floats;
    H.threshold = (float)threshold;
    std::vector<unsigned char> blob(sizeof(SceneHeader), 0);
    H.off_body = append(blob, bodies);
    H.off_joint = append(blob, joints);
    H.off_geom = append(blob, geoms);
    H.off_const = append(blob, consts);
    H.off_rec = append(blob, recs);
    H.off_pair = append(blob, pairs);
    H.off_cull = append(blob, cull);
    H.off_group = append(blob, groups);
    H.n_group = (int)groups.size();
    std::vector<uint32_t> wmask(pairs.size() / 32 * 4, 0u);
    for (size_t i = 0; i < pairs.size(); i++) {
        const int cls = pairs[i].cls;
        if (cls > PC_MPR) continue;   // dummies
        const int list = cls < PC_BOX_BOX ? 0 : (cls == PC_BOX_BOX ? 1 : 2);
        wmask[i / 32 * 4 + list] |= 1u << (i % 32);
    }
    H.off_wmask = append(blob, wmask);
    H.off_real = append(blob, real);
    H.n_real = (int)real.size();
    H.n_pruned = n_pruned;
    std::vector<float> hull(3 * (size_t)d->nmeshvert);
    for (size_t k = 0; k < hull.size(); k++) hull[k] = (float)d->mesh_vert[k];
    H.off_hull = append(blob, hull);
    H.n_hull_vert = d->nmeshvert;
    for (size_t r = 0; r < recs.size(); r++) {   // mesh records: locate their hull relative to the record itself
        if (recs[r].kind != K_MESH) continue;
        const int me = d->geom_dataid[recs[r].geom_id];
        GeomRec *Rr = reinterpret_cast<GeomRec *>(blob.data() + H.off_rec) + r;
        const int rel = (H.off_hull + 12 * d->mesh_vertadr[me]) - (H.off_rec + (int)(r * sizeof(GeomRec)));
        const int cnt = d->mesh_vertnum[me];
        memcpy(&Rr->sx, &rel, 4);
        memcpy(&Rr->sy, &cnt, 4);
    }
    while (blob.size() % 16) blob.push_back(0);
    H.blob_bytes = (int)blob.size();
    memcpy(blob.data(), &H, sizeof(H));
    out.hdr = H;
    out.blob.swap(blob);
}

}  // namespace mopa
