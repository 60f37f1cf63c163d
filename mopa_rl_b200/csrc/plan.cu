// RRT-Connect over the active joints, sm_100a: one warp per planning problem.
//
// Replaces KinematicPlanner::plan (motion_planners/KinematicPlanner.cpp:125-251) and the OMPL
// pieces it drives (RRTConnect::solve/growTree, DiscreteMotionValidator::checkMotion,
// CompoundStateSpace distance/interpolate).  Per main-loop iteration a
// warp: draws the sample (counter-based RNG), finds the nearest tree node with a strided scan
// + lexicographic (distance, index) shuffle reduction, builds the extension state, and
// validates the new state and every interior state of the edge at once - lane k runs the
// forward kinematics of state k, then all 32 lanes sweep the (state, candidate pair) items.
// Trees live in global memory (L2 resident); the scene tables and per-warp frames in shared.
// Termination is an iteration cap (the reference's wall-clock `timelimit` is not reproducible).
#include <cuda_runtime.h>

#include <stdexcept>

#include "planner_state.h"
#include "validity_kernel.cuh"

namespace mopa {

void build_space(const mopa_model_desc *d, const int32_t *passive, int n_passive, double range, double resolution,
                 uint64_t seed, Space &out) {
    out = Space();
    out.nq = d->nq;
    out.range = (float)range;
    out.resolution = (float)resolution;
    out.seed = seed;
    for (int j = 0; j < d->njnt; j++) {
        int adr = d->jnt_qposadr[j];
        bool is_passive = false;
        for (int k = 0; k < n_passive; k++)
            if (passive[k] == adr) is_passive = true;
        if (is_passive) continue;
        int t = d->jnt_type[j];
        if (t == MOPA_JNT_FREE || t == MOPA_JNT_BALL)
            throw std::runtime_error("free/ball joints cannot be active planner joints (pass their qpos indices as passive)");
        out.active_qadr.push_back(adr);
        if (t == MOPA_JNT_HINGE && !d->jnt_limited[j]) {
            out.is_so2.push_back(1);
            out.lo.push_back(-3.14159265358979323846f);
            out.hi.push_back(3.14159265358979323846f);
        } else {
            out.is_so2.push_back(0);
            out.lo.push_back((float)d->jnt_range[2 * j]);
            out.hi.push_back((float)d->jnt_range[2 * j + 1]);
        }
    }
    out.n_active = (int)out.active_qadr.size();
    // the reference throws when the joint dimensions do not add up (mujoco_ompl_interface.cpp:268-272)
    if (out.n_active != d->nq - n_passive) throw std::runtime_error("Total joint dimensions are not equal to nq - size(passive_joints)");
    if (out.n_active > PLAN_MAXD) throw std::runtime_error("more than 8 active joints are not supported");
}

struct SpaceDev {
    int nd, nq;
    int adr[PLAN_MAXD];
    int so2[PLAN_MAXD];
    float lo[PLAN_MAXD], hi[PLAN_MAXD], seg[PLAN_MAXD];
    float range;
    unsigned long long seed;
};

#define PI_F 3.14159265358979323846f

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
__device__ __forceinline__ float urand(unsigned long long seed, unsigned long long key, unsigned it, unsigned dim) {
    unsigned long long x = mix64(seed ^ (key * 0x9E3779B97F4A7C15ULL));
    x = mix64(x + ((((unsigned long long)it) << 8) | dim) * 0xD1342543DE82EF95ULL);
    return (float)(unsigned)(x >> 40) * 5.9604644775390625e-08f;
}

struct St { float v[PLAN_MAXD]; };

__device__ __forceinline__ float dist1(const SpaceDev &sp, int j, float a, float b) {
    float d = fabsf(a - b);
    if (sp.so2[j] && d > PI_F) d = 2.0f * PI_F - d;
    return d;
}
__device__ __forceinline__ float distance(const SpaceDev &sp, const St &a, const St &b) {
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < PLAN_MAXD; j++)
        if (j < sp.nd) d = d + dist1(sp, j, a.v[j], b.v[j]);
    return d;
}
__device__ __forceinline__ St interpolate(const SpaceDev &sp, const St &a, const St &b, float t) {
    St o;
#pragma unroll
    for (int j = 0; j < PLAN_MAXD; j++) {
        o.v[j] = 0.f;
        if (j >= sp.nd) continue;
        float diff = b.v[j] - a.v[j];
        if (!sp.so2[j] || fabsf(diff) <= PI_F) o.v[j] = fmaf(diff, t, a.v[j]);
        else {
            if (diff > 0) diff = 2.0f * PI_F - diff; else diff = -2.0f * PI_F - diff;
            float v = fmaf(-diff, t, a.v[j]);
            if (v > PI_F) v -= 2.0f * PI_F; else if (v < -PI_F) v += 2.0f * PI_F;
            o.v[j] = v;
        }
    }
    return o;
}
__device__ __forceinline__ int seg_count(const SpaceDev &sp, const St &a, const St &b) {
    int n = 0;
#pragma unroll
    for (int j = 0; j < PLAN_MAXD; j++)
        if (j < sp.nd) n = max(n, (int)ceilf(dist1(sp, j, a.v[j], b.v[j]) / sp.seg[j]));
    return n;
}

struct SmemRowQ {
    const float *row;
    __device__ __forceinline__ float operator()(int i) const { return row[i]; }
};

constexpr int PW_STATES = 8;     // states validated per sweep
constexpr int PLAN_WARPS = 8;    // problems per CTA

struct WarpCtx {
    SceneView S;
    float *rows;     // [PW_STATES][row_pad]
    float *frames;   // [frame_floats][PW_STATES]
    int row_pad;
    int lane;
};

// All `count` (<= PW_STATES) states valid?  Lane k < count holds state k in `mine`.
template <bool MESH>
__device__ __forceinline__ bool states_all_valid(const WarpCtx &W, const SpaceDev &sp, const float *baseq, const St &mine, int count) {
    const int lane = W.lane;
    if (lane < count) {
        float *row = W.rows + lane * W.row_pad;
        for (int i = 0; i < sp.nq; i++) row[i] = baseq[i];
#pragma unroll
        for (int j = 0; j < PLAN_MAXD; j++)
            if (j < sp.nd) row[sp.adr[j]] = mine.v[j];
        SmemRowQ rq{row};
        fk_state(W.S, rq, W.frames, PW_STATES, lane);
    }
    __syncwarp();
    const int npair = W.S.H->n_real;   // the non-dummy entries of the pair table
    const int total = count * npair;
    const float thr = W.S.H->threshold;
    bool bad = false;
    for (int base = 0; base < total; base += 32) {
        const int i = base + lane;
        if (i < total) {
            const int k = i / npair, p = W.S.real[i - k * npair];
            const PairRec pr = W.S.pairs[p];
            const bool survive = cull_survives(pr, W.S.cull[p], W.frames, PW_STATES, k);
            if (survive && pr.cls <= PC_MPR) {
                Geom a, b;
                load_geom<MESH>(a, W.S.recs[pr.ga], W.frames, PW_STATES, k);
                load_geom<MESH>(b, W.S.recs[pr.gb], W.frames, PW_STATES, k);
                float dist = pr.cls >= PC_BOX_BOX ? heavy_dist<MESH>(pr.cls, a, b) : cheap_dist<MESH>(pr.cls, a, b, thr);
                if (dist <= thr) bad = true;
            }
        }
        if (__any_sync(0xffffffffu, bad)) { bad = true; break; }
    }
    __syncwarp();
    return !bad;
}

enum { G_TRAPPED = 0, G_ADVANCED = 1, G_REACHED = 2 };

struct Tree {
    float *x;      // [max_nodes][PLAN_MAXD]
    int *parent;   // [max_nodes]
    int n;
};

__device__ __forceinline__ St load_state(const float *x, int i) {
    St s;
    const float4 *p = reinterpret_cast<const float4 *>(x + (size_t)i * PLAN_MAXD);
    float4 a = p[0], b = p[1];
    s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.v[3] = a.w; s.v[4] = b.x; s.v[5] = b.y; s.v[6] = b.z; s.v[7] = b.w;
    return s;
}

__device__ __forceinline__ int nearest(const SpaceDev &sp, const Tree &t, const St &x, int lane) {
    float bd = 3.0e38f;
    int bi = 0x7fffffff;
    for (int i = lane; i < t.n; i += 32) {
        float d = distance(sp, load_state(t.x, i), x);
        if (d < bd) { bd = d; bi = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, bd, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
    }
    return bi;
}

template <bool MESH>
__device__ __forceinline__ int grow(const WarpCtx &W, const SpaceDev &sp, const float *baseq, Tree &t, bool is_start,
                                    const St &target, St &xstate, int &added, int max_nodes) {
    const int lane = W.lane;
    const int ni = nearest(sp, t, target, lane);
    const St ns = load_state(t.x, ni);
    const float d = distance(sp, ns, target);
    St dstate = target;
    bool reach = true;
    if (d > sp.range) {
        dstate = interpolate(sp, ns, target, sp.range / d);
        bool same = true;
#pragma unroll
        for (int j = 0; j < PLAN_MAXD; j++)
            if (j < sp.nd && dstate.v[j] != ns.v[j]) same = false;
        if (same) return G_TRAPPED;
        reach = false;
    }
    // states to validate: dstate and the interior points of the edge (direction matters for rounding)
    const St s1 = is_start ? ns : dstate, s2 = is_start ? dstate : ns;
    const int nd = seg_count(sp, s1, s2);
    const int nstates = 1 + max(nd - 1, 0);
    for (int base = 0; base < nstates; base += PW_STATES) {
        const int cnt = min(PW_STATES, nstates - base);
        St mine = dstate;
        const int idx = base + lane;  // 0: dstate, m>=1: interior point m
        if (lane < cnt && idx >= 1) mine = interpolate(sp, s1, s2, (float)idx / (float)nd);
        if (!states_all_valid<MESH>(W, sp, baseq, mine, cnt)) return G_TRAPPED;
    }
    if (t.n >= max_nodes) return G_TRAPPED;
    if (lane == 0) {
        float4 *p = reinterpret_cast<float4 *>(t.x + (size_t)t.n * PLAN_MAXD);
        p[0] = make_float4(dstate.v[0], dstate.v[1], dstate.v[2], dstate.v[3]);
        p[1] = make_float4(dstate.v[4], dstate.v[5], dstate.v[6], dstate.v[7]);
        t.parent[t.n] = ni;
    }
    __syncwarp();
    added = t.n;
    t.n++;
    xstate = dstate;
    return reach ? G_REACHED : G_ADVANCED;
}

template <bool MESH>
__global__ void __launch_bounds__(PLAN_WARPS * 32)
plan_kernel(const unsigned char *__restrict__ blob_g, int blob_bytes, SpaceDev sp, const float *__restrict__ start,
            const float *__restrict__ goal, int row_stride, const unsigned long long *__restrict__ keys, int n, int max_iter,
            float *__restrict__ tree_x, int *__restrict__ tree_parent, int max_nodes, float *__restrict__ path,
            int *__restrict__ node_ids, int max_path, int *__restrict__ path_len, int *__restrict__ status_out,
            int *__restrict__ iters_out, int *__restrict__ nodes_out, const int *__restrict__ d_n, int only_failed,
            unsigned long long key_xor) {
    extern __shared__ __align__(16) unsigned char smem[];
    if (d_n) { const int m = *d_n; if (m < n) n = m; if (n <= 0) return; }   // problem count produced on the device
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < blob_bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(blob_g)[i];
    __syncthreads();
    WarpCtx W;
    W.S = view_scene(smem);
    W.lane = lane;
    W.row_pad = sp.nq + 1;
    const int ff = W.S.H->frame_floats;
    const size_t per_warp = (size_t)PW_STATES * W.row_pad + (size_t)ff * PW_STATES + sp.nq;
    float *wbase = reinterpret_cast<float *>(smem + blob_bytes) + per_warp * warp;
    W.rows = wbase;
    W.frames = wbase + PW_STATES * W.row_pad;
    float *baseq = W.frames + (size_t)ff * PW_STATES;

    const int cta_warps = blockDim.x >> 5;   // 8 for stand-alone batches; 1 when the CTAs are meant to co-reside with the env-step kernel
    for (int prob = blockIdx.x * cta_warps + warp; prob < n; prob += gridDim.x * cta_warps) {
        if (only_failed && status_out[prob] == 0) continue;   // retry pass: problems an earlier planner already solved keep their path
        const float *srow = start + (size_t)prob * row_stride, *grow_ = goal + (size_t)prob * row_stride;
        for (int i = lane; i < sp.nq; i += 32) baseq[i] = srow[i];  // passive dims frozen at the start values
        __syncwarp();
        St s, g;
#pragma unroll
        for (int j = 0; j < PLAN_MAXD; j++) {
            s.v[j] = j < sp.nd ? srow[sp.adr[j]] : 0.f;
            g.v[j] = j < sp.nd ? grow_[sp.adr[j]] : 0.f;
        }
        const unsigned long long key = keys[prob] ^ key_xor;
        int status = MOPA_PLAN_NOT_EXACT_, it = 0, sm = -1, gm = -1;
        Tree T[2];
        for (int k = 0; k < 2; k++) {
            T[k].x = tree_x + ((size_t)prob * 2 + k) * max_nodes * PLAN_MAXD;
            T[k].parent = tree_parent + ((size_t)prob * 2 + k) * max_nodes;
            T[k].n = 0;
        }
        bool ok = true;
        if (!states_all_valid<MESH>(W, sp, baseq, g, 1)) { status = MOPA_PLAN_INVALID_GOAL_; ok = false; }
        if (ok) {
            bool inb = true;
#pragma unroll
            for (int j = 0; j < PLAN_MAXD; j++)
                if (j < sp.nd && !sp.so2[j]) {
                    if (s.v[j] > sp.hi[j] || s.v[j] < sp.lo[j]) inb = false;
                    if (g.v[j] > sp.hi[j] || g.v[j] < sp.lo[j]) inb = false;
                }
            if (!inb || !states_all_valid<MESH>(W, sp, baseq, s, 1)) ok = false;
        }
        if (ok) {
            if (lane == 0) {
                for (int k = 0; k < 2; k++) {
                    const St &r = k ? g : s;
                    float4 *p = reinterpret_cast<float4 *>(T[k].x);
                    p[0] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
                    p[1] = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
                    T[k].parent[0] = -1;
                }
            }
            __syncwarp();
            T[0].n = T[1].n = 1;
            bool start_tree = true;
            for (it = 0; it < max_iter; it++) {
                const int ti = start_tree ? 0 : 1;
                const bool is_start = start_tree;
                start_tree = !start_tree;
                const int oi = start_tree ? 0 : 1;
                St rstate, xstate;
#pragma unroll
                for (int j = 0; j < PLAN_MAXD; j++)
                    rstate.v[j] = j < sp.nd ? fmaf(sp.hi[j] - sp.lo[j], urand(sp.seed, key, (unsigned)it, (unsigned)j), sp.lo[j]) : 0.f;
                int added = -1, oadded = -1;
                int gs = grow<MESH>(W, sp, baseq, T[ti], is_start, rstate, xstate, added, max_nodes);
                if (gs == G_TRAPPED) continue;
                rstate = xstate;
                int gsc = grow<MESH>(W, sp, baseq, T[oi], start_tree, rstate, xstate, oadded, max_nodes);
                while (gsc == G_ADVANCED) gsc = grow<MESH>(W, sp, baseq, T[oi], start_tree, rstate, xstate, oadded, max_nodes);
                if (gsc == G_REACHED) {
                    sm = start_tree ? oadded : added;
                    gm = start_tree ? added : oadded;
                    status = 0;
                    it++;
                    break;
                }
            }
        }
        int cnt = 0;
        if (status == 0) {
            if (T[0].parent[sm] >= 0) sm = T[0].parent[sm]; else gm = T[1].parent[gm];
            int n1 = 0;
            for (int i = sm; i >= 0; i = T[0].parent[i]) n1++;
            cnt = n1;
            for (int i = gm; i >= 0; i = T[1].parent[i]) cnt++;
            if (cnt > max_path) { status = MOPA_PLAN_NOT_EXACT_; cnt = 0; }
            else {
                float *prow = path + (size_t)prob * max_path * row_stride;
                int *ids = node_ids + (size_t)prob * max_path;
                int r = n1 - 1;
                for (int i = sm; i >= 0; i = T[0].parent[i], r--) {
                    for (int c = lane; c < sp.nq; c += 32) prow[(size_t)r * row_stride + c] = baseq[c];
                    __syncwarp();
                    if (lane < sp.nd) prow[(size_t)r * row_stride + sp.adr[lane]] = T[0].x[(size_t)i * PLAN_MAXD + lane];
                    if (lane == 0) ids[r] = i;
                }
                r = n1;
                for (int i = gm; i >= 0; i = T[1].parent[i], r++) {
                    for (int c = lane; c < sp.nq; c += 32) prow[(size_t)r * row_stride + c] = baseq[c];
                    __syncwarp();
                    if (lane < sp.nd) prow[(size_t)r * row_stride + sp.adr[lane]] = T[1].x[(size_t)i * PLAN_MAXD + lane];
                    if (lane == 0) ids[r] = i | (1 << 30);
                }
            }
        }
        if (lane == 0) {
            status_out[prob] = status;
            path_len[prob] = cnt;
            if (iters_out) iters_out[prob] = it;
            if (nodes_out) { nodes_out[2 * prob] = T[0].n; nodes_out[2 * prob + 1] = T[1].n; }
        }
        __syncwarp();
    }
}

struct PlanBuffers {
    size_t cap = 0;
    int max_nodes = 0;
    float *tree_x = nullptr;
    int *tree_parent = nullptr;
    // staging for the host entry point
    size_t hcap = 0;
    int hmax_path = 0;
    float *d_start = nullptr, *d_goal = nullptr, *d_path = nullptr;
    unsigned long long *d_keys = nullptr;
    int *d_ids = nullptr, *d_len = nullptr, *d_status = nullptr, *d_iters = nullptr, *d_nodes = nullptr;
};

void free_plan_buffers(mopa_planner *p) {
    PlanBuffers *b = (PlanBuffers *)p->plan_buffers;
    if (!b) return;
    cudaFree(b->tree_x); cudaFree(b->tree_parent); cudaFree(b->d_start); cudaFree(b->d_goal); cudaFree(b->d_path);
    cudaFree(b->d_keys); cudaFree(b->d_ids); cudaFree(b->d_len); cudaFree(b->d_status); cudaFree(b->d_iters); cudaFree(b->d_nodes);
    delete b;
    p->plan_buffers = nullptr;
}

static cudaError_t ensure_trees(mopa_planner *p, size_t n, int max_nodes) {
    if (!p->plan_buffers) p->plan_buffers = new PlanBuffers();
    PlanBuffers *b = (PlanBuffers *)p->plan_buffers;
    if (n <= b->cap && max_nodes == b->max_nodes) return cudaSuccess;
    cudaFree(b->tree_x); cudaFree(b->tree_parent);
    b->tree_x = nullptr; b->tree_parent = nullptr; b->cap = 0;
    size_t cap = n < 64 ? 64 : n;
    cudaError_t e = cudaMalloc(&b->tree_x, cap * 2 * (size_t)max_nodes * PLAN_MAXD * sizeof(float));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&b->tree_parent, cap * 2 * (size_t)max_nodes * sizeof(int));
    if (e != cudaSuccess) return e;
    b->cap = cap;
    b->max_nodes = max_nodes;
    return cudaSuccess;
}

cudaError_t launch_plan(mopa_planner *p, const float *d_start, const float *d_goal, int row_stride, const unsigned long long *d_keys,
                        int n, int max_iter, float *d_path, int *d_node_ids, int max_path, int *d_path_len, int *d_status,
                        int *d_iters, int *d_nodes, cudaStream_t stream, const int *d_n, int cta_warps, float range_override, int only_failed,
                        unsigned long long key_xor) {
    if (n <= 0) return cudaSuccess;
    if (cta_warps < 1 || cta_warps > PLAN_WARPS) cta_warps = PLAN_WARPS;
    cudaError_t e = ensure_trees(p, (size_t)n, p->max_nodes);
    if (e != cudaSuccess) return e;
    PlanBuffers *b = (PlanBuffers *)p->plan_buffers;
    SpaceDev sp;
    memset(&sp, 0, sizeof(sp));
    sp.nd = p->space.n_active; sp.nq = p->space.nq; sp.range = range_override > 0.f ? range_override : p->space.range; sp.seed = p->space.seed;
    for (int j = 0; j < sp.nd; j++) {
        sp.adr[j] = p->space.active_qadr[j]; sp.so2[j] = p->space.is_so2[j];
        sp.lo[j] = p->space.lo[j]; sp.hi[j] = p->space.hi[j];
        float ext = sp.so2[j] ? PI_F : (sp.hi[j] - sp.lo[j]);
        sp.seg[j] = p->space.resolution * ext;
    }
    const SceneHeader &H = p->scene.hdr;
    size_t per_warp = ((size_t)PW_STATES * (H.nq + 1) + (size_t)H.frame_floats * PW_STATES + H.nq) * sizeof(float);
    size_t smem = (size_t)H.blob_bytes + per_warp * cta_warps + 16;
    static bool attr_set[2] = {false, false};
    const int mesh = H.n_hull_vert > 0;   // scenes with mesh colliders: instantiation with the hull support function
    auto kern = mesh ? plan_kernel<true> : plan_kernel<false>;
    if (!attr_set[mesh]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[mesh] = true;
    }
    int grid = (n + cta_warps - 1) / cta_warps;
    int max_grid = cta_warps == 1 ? p->sm_count * 2 : p->sm_count * 4;
    if (grid > max_grid) grid = max_grid;
    kern<<<grid, cta_warps * 32, smem, stream>>>(p->d_blob, H.blob_bytes, sp, d_start, d_goal, row_stride, d_keys, n, max_iter,
                                                        b->tree_x, b->tree_parent, b->max_nodes, d_path, d_node_ids, max_path,
                                                        d_path_len, d_status, d_iters, d_nodes, d_n, only_failed, key_xor);
    return cudaGetLastError();
}

// host-buffer planning: staging + launch + copy back only the rows that were written
int plan_host(mopa_planner *p, const double *start, const double *goal, const uint64_t *keys, int n, int max_iter, double *path,
              int32_t *node_ids, int max_path, int32_t *path_len, int32_t *status, int32_t *iters, std::string &err) {
    if (!p->plan_buffers) p->plan_buffers = new PlanBuffers();
    PlanBuffers *b = (PlanBuffers *)p->plan_buffers;
    const int nq = p->scene.hdr.nq, row = p->scene.hdr.nq4 * 4;
#define PH_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { err = std::string(#x) + ": " + cudaGetErrorString(e_); return -2; } } while (0)
    if ((size_t)n > b->hcap || max_path != b->hmax_path) {
        cudaFree(b->d_start); cudaFree(b->d_goal); cudaFree(b->d_path); cudaFree(b->d_keys); cudaFree(b->d_ids);
        cudaFree(b->d_len); cudaFree(b->d_status); cudaFree(b->d_iters); cudaFree(b->d_nodes);
        b->hcap = 0;
        size_t cap = n < 16 ? 16 : n;
        PH_TRY(cudaMalloc(&b->d_start, cap * row * sizeof(float)));
        PH_TRY(cudaMalloc(&b->d_goal, cap * row * sizeof(float)));
        PH_TRY(cudaMalloc(&b->d_path, cap * (size_t)max_path * row * sizeof(float)));
        PH_TRY(cudaMalloc(&b->d_keys, cap * sizeof(unsigned long long)));
        PH_TRY(cudaMalloc(&b->d_ids, cap * (size_t)max_path * sizeof(int)));
        PH_TRY(cudaMalloc(&b->d_len, cap * sizeof(int)));
        PH_TRY(cudaMalloc(&b->d_status, cap * sizeof(int)));
        PH_TRY(cudaMalloc(&b->d_iters, cap * sizeof(int)));
        PH_TRY(cudaMalloc(&b->d_nodes, cap * 2 * sizeof(int)));
        b->hcap = cap;
        b->hmax_path = max_path;
    }
    std::vector<float> hs((size_t)n * row, 0.f), hg((size_t)n * row, 0.f);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < nq; k++) {
            hs[(size_t)i * row + k] = (float)start[(size_t)i * nq + k];
            hg[(size_t)i * row + k] = (float)goal[(size_t)i * nq + k];
        }
    cudaStream_t st = p->stream;
    PH_TRY(cudaMemcpyAsync(b->d_start, hs.data(), hs.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    PH_TRY(cudaMemcpyAsync(b->d_goal, hg.data(), hg.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    PH_TRY(cudaMemcpyAsync(b->d_keys, keys, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PH_TRY(launch_plan(p, b->d_start, b->d_goal, row, b->d_keys, n, max_iter, b->d_path, b->d_ids, max_path, b->d_len, b->d_status,
                       b->d_iters, b->d_nodes, st));
    PH_TRY(cudaMemcpyAsync(path_len, b->d_len, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    PH_TRY(cudaMemcpyAsync(status, b->d_status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (iters) PH_TRY(cudaMemcpyAsync(iters, b->d_iters, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    PH_TRY(cudaStreamSynchronize(st));
    std::vector<float> hp;
    for (int i = 0; i < n; i++) {
        int len = path_len[i];
        if (len <= 0) continue;
        hp.resize((size_t)len * row);
        PH_TRY(cudaMemcpyAsync(hp.data(), b->d_path + (size_t)i * max_path * row, hp.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (node_ids) PH_TRY(cudaMemcpyAsync(node_ids + (size_t)i * max_path, b->d_ids + (size_t)i * max_path, (size_t)len * sizeof(int), cudaMemcpyDeviceToHost, st));
        PH_TRY(cudaStreamSynchronize(st));
        for (int r = 0; r < len; r++)
            for (int k = 0; k < nq; k++) path[((size_t)i * max_path + r) * nq + k] = (double)hp[(size_t)r * row + k];
    }
#undef PH_TRY
    return 0;
}

}  // namespace mopa
