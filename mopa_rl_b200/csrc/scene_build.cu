// Host-side scene compiler: mjModel-style arrays -> device tables for the validity kernel.
//
// Covers what the reference does at planner construction
// (KinematicPlanner.cpp:42-120: load model, build the state space over active joints,
// install MujocoStateValidityChecker with `ignored_contacts` and `contact_threshold`) plus
// the query-independent half of MuJoCo's collision pipeline: the
// candidate-pair filters and the world frames of world-welded geoms.
#include <algorithm>
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
        if (X.list < 3) pb.push_back(X);   // pairs of classes the narrow phase does not know never collide (PC_NONE)
    }
    // kernel order: runs of (cull kind, anchor); every run is padded to a multiple of four entries with dummy pairs that never
    // survive the cull (the sweep tests four entries per trip without bounds checks)
    std::stable_sort(pb.begin(), pb.end(), [](const PairBuild &a, const PairBuild &b) {
        if (a.P.ckind != b.P.ckind) return a.P.ckind < b.P.ckind;
        if (a.P.anchor_slot != b.P.anchor_slot) return a.P.anchor_slot < b.P.anchor_slot;
        return a.P.canon < b.P.canon;
    });
    std::vector<PairRec> pairs;
    std::vector<CullEntry> cull;
    std::vector<CullGroup> groups;
    auto pad_group = [&]() {
        if (groups.empty() || groups.back().kind == CK_NONE) return;
        while (groups.back().count % 4) {
            PairRec D;
            memset(&D, 0, sizeof(D));
            D.cls = PC_NONE; D.ckind = (uint8_t)groups.back().kind; D.anchor_slot = groups.back().anchor_slot; D.partner_slot = groups.back().anchor_slot;
            D.canon = 0xFFFF;
            CullEntry E{0.0f, 0.0f, 0.0f, -1.0f};   // |d|^2 > -1 and 0 > -1: culled by every kind of test
            if (groups.back().kind == CK_SPHERE_MOVING) { const int off = groups.back().anchor_slot; memcpy(&E.x, &off, 4); }
            pairs.push_back(D); cull.push_back(E);
            groups.back().count++;
        }
    };
    for (size_t i = 0; i < pb.size(); i++) {
        const bool fresh = i == 0 || pb[i].P.ckind != pb[i - 1].P.ckind || pb[i].P.anchor_slot != pb[i - 1].P.anchor_slot;
        if (fresh) { pad_group(); groups.push_back(CullGroup{pb[i].P.anchor_slot, pb[i].P.ckind, 0, 0}); }
        pairs.push_back(pb[i].P);
        cull.push_back(pb[i].E);
        groups.back().count++;
    }
    pad_group();
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
