// mopa_is_valid_batch: batched state-validity checks on sm_100a (replaces MujocoStateValidityChecker::isValid,
// motion_planners/src/mujoco_ompl_interface.cpp:909-978).  See the kernel comment below for the mapping.
// HBM traffic is the qpos row in and one 32-bit word out per query.
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <utility>

#include "validity_kernel.cuh"

namespace mopa {

struct RowQ {
    const float *row;
    __device__ __forceinline__ float operator()(int i) const { return __ldg(row + i); }
};

// ---- the flattened checker (round 2).  Round 1 swept the candidate pairs with one thread per query and evaluated the
// narrow phase where it stood: ~25 of the 241 (query, pair) items survive the bounding-sphere cull (14 of them plane pairs
// that had no cull at all), 0.8 per query end in the portal refinement whose iteration count has a heavy tail, and the
// warp waited for its slowest lane everywhere (ncu: 11.9 of 32 threads per instruction).  Now every kind of work runs in
// the shape that suits it:
//   phase A  thread = query   forward kinematics (serial chain, in registers); world frames of the moving geoms go to shared
//                             memory ([frame float][query]: conflict-free for thread = query access)
//   phase B  thread = query   a pure cull sweep over the candidate pairs, no narrow phase inside: pair records are broadcast
//                             reads, the control flow is uniform (bounding spheres; plane pairs by centre height - bounding
//                             radius), survivors are kept as a bit mask in registers and then appended to one of three
//                             CTA-wide work lists (analytic pairs / box-box / portal-refinement candidates)
//   phase C  thread = item    the CTA drains the lists one at a time: the lanes of a warp run the same routine on different
//                             items.  Portal-refinement candidates first pass the conservative segment / slab pre-test;
//                             what is left is queued (with the two world frames) in a per-CTA slab in global memory
//   phase D  lane = item, refilled   once ~8 items per thread are queued (or the CTA runs out of tiles) the refinement runs
//                             as a state machine, one support evaluation per trip; a lane whose item is finished takes the
//                             next one, so the tail of long-running items no longer idles the other 31 lanes
// During the kernel out[] holds the raw verdict (smallest canonical index of an offending pair, or ~0); the CTA converts
// its own rows to the result-word format at the end.  Cull and narrow-phase arithmetic are round 1's, on the same inputs:
// the result words are unchanged (bit-identical to the f32 oracle).  A full list / queue makes the pushing lane evaluate
// its item in place.
enum { WL_CHEAP = 0, WL_BOX = 1, WL_MPR = 2, WL_COUNT = 3 };
#ifdef MOPA_VK_STATS   // diagnostics build (tools/vk_stats.py): per pair class, pairs tested / cull survivors / offending / refined
__device__ unsigned long long g_vk_stats[16][4];
#define VK_STAT(cls, k) atomicAdd(&g_vk_stats[cls][k], 1ULL)
#else
#define VK_STAT(cls, k)
#endif

struct MprItem { int q, p; float ca[3], Ra[9], cb[3], Rb[9]; };   // 104 bytes: global query row, pair, the two world frames

// one (query, pair) item of the analytic / box-box classes: distance against the threshold
// (WITH_MPR = false: the caller only passes analytic and box-box pairs - phase C; keeps the portal routine out of its loop body)
template <bool MESH, bool WITH_MPR = true>
__device__ __forceinline__ void eval_item(const SceneView &S, const float *frames, int stride, int ql, const PairRec &pr, float thr, uint32_t *res_q) {
    Geom a, b;
    load_geom<MESH>(a, S.recs[pr.ga], frames, stride, ql);
    load_geom<MESH>(b, S.recs[pr.gb], frames, stride, ql);
    float dist;
    if (WITH_MPR) dist = pr.cls >= PC_BOX_BOX ? heavy_dist<MESH>(pr.cls, a, b) : cheap_dist<MESH>(pr.cls, a, b, thr);
    else dist = pr.cls == PC_BOX_BOX ? box_box(a, b) : cheap_dist<MESH>(pr.cls, a, b, thr);
    if (dist <= thr) { atomicMin(res_q, (uint32_t)pr.canon); VK_STAT(pr.cls, 2); }
}
// the same for any class, out of line: a work list / queue that is full makes the pushing lane evaluate its item in place (rare)
template <bool MESH>
__device__ __noinline__ void eval_item_overflow(const unsigned char *blob, const float *frames, int stride, int ql, int p, float thr, uint32_t *res_q) {
    const SceneView S = view_scene(blob);
    eval_item<MESH>(S, frames, stride, ql, S.pairs[p], thr, res_q);
}

template <int NQ, bool MESH>
__global__ void __launch_bounds__(NQ, NQ <= 128 ? 3 : 1) is_valid_kernel(const unsigned char *__restrict__ blob_g, int blob_bytes,
                                                      const float *__restrict__ qpos, int row_stride, int n,
                                                      uint32_t *__restrict__ out, int exact, const int *__restrict__ d_n, int d_n_mult,
                                                      int ffs, int cap_cheap, int cap_box, int cap_mpr, MprItem *__restrict__ mq_all,
                                                      int mq_cap, int mq_run) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    if (d_n) {   // row count produced on the device (no host round trip): n is the capacity
        const long long m = (long long)(*d_n) * d_n_mult;
        if (m < n) n = (int)m;
        if (n <= 0) return;
    }
    for (int i = tid; i < blob_bytes / 16; i += NQ) reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(blob_g)[i];
    __syncthreads();
    const SceneView S = view_scene(smem);
    float *frames = reinterpret_cast<float *>(smem + blob_bytes);         // [ffs][NQ]
    uint32_t *res = reinterpret_cast<uint32_t *>(frames + (size_t)ffs * NQ);
    int *wcount = reinterpret_cast<int *>(res + NQ);                        // [0..2] work lists, [4] queue length, [5] queue cursor, [8..10] list offsets, [12..14] capacities
    uint32_t *wl0 = reinterpret_cast<uint32_t *>(wcount + 48);              // the three lists, back to back; wcount[16..47]: key histogram / offsets of the queue sort
    auto wl_base = [&](int list) { return wl0 + wcount[8 + list]; };
    auto wl_cap = [&](int list) { return wcount[12 + list]; };
    MprItem *mq = mq_all + (size_t)blockIdx.x * mq_cap;
    if (tid == 0) {
        wcount[4] = 0; wcount[5] = 0;
        wcount[8] = 0; wcount[9] = cap_cheap; wcount[10] = cap_cheap + cap_box;
        wcount[12] = cap_cheap; wcount[13] = cap_box; wcount[14] = cap_mpr;
    }
    __syncthreads();
    const float thr = S.H->threshold;
    const int npair = S.H->n_pair;
    const int ntile = (n + NQ - 1) / NQ;

    // ---- phase D: the queued portal refinements, lanes refilled from the queue
    auto run_queue = [&]() {
        const int cnt = min(wcount[4], mq_cap);
        __syncthreads();
        if (cnt > 0) {
            // group the queue by the (kind, kind) combination of its pairs: a counting sort of item indices into the idle
            // work-list memory, so that neighbouring lanes - which take neighbouring positions - run the same support routines
            uint16_t *order = reinterpret_cast<uint16_t *>(wl0);
            int *hist = wcount + 16;
            if (tid < 32) hist[tid] = 0;
            __syncthreads();
            for (int i = tid; i < cnt; i += NQ) atomicAdd(&hist[S.pairs[mq[i].p].mkey], 1);
            __syncthreads();
            if (tid == 0) { int run = 0; for (int k = 0; k < 16; k++) { const int c = hist[k]; hist[16 + k] = run; run += c; } }
            __syncthreads();
            for (int i = tid; i < cnt; i += NQ) order[atomicAdd(&hist[16 + S.pairs[mq[i].p].mkey], 1)] = (uint16_t)i;
            __syncthreads();
            MprSM m;
            Geom a, b;
            int iq = 0;
            uint32_t canon = 0;
            bool active = false, exhausted = false;
            for (;;) {
                const unsigned act = __ballot_sync(0xffffffffu, active);
                if (__popc(act) <= 24 && !active && !exhausted) {   // refill once a quarter of the warp is idle
                    const int idx = atomicAdd(&wcount[5], 1);
                    if (idx >= cnt) exhausted = true;
                    else {
                        const MprItem &it = mq[order[idx]];
                        iq = it.q;
                        if (exact || out[iq] == 0xFFFFFFFFu) {   // fast mode: nothing to learn about a state already known to be invalid
                            const PairRec pr = S.pairs[it.p];
                            canon = pr.canon;
                            const GeomRec &ra = S.recs[pr.ga], &rb = S.recs[pr.gb];
                            a.kind = ra.kind; a.size = V3{ra.sx, ra.sy, ra.sz}; a.hull = nullptr; a.nhull = 0;
                            b.kind = rb.kind; b.size = V3{rb.sx, rb.sy, rb.sz}; b.hull = nullptr; b.nhull = 0;
                            if (MESH && ra.kind == K_MESH) { a.hull = reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(&ra) + __float_as_int(ra.sx)); a.nhull = __float_as_int(ra.sy); a.size.z = ra.rbound; }
                            if (MESH && rb.kind == K_MESH) { b.hull = reinterpret_cast<const float *>(reinterpret_cast<const unsigned char *>(&rb) + __float_as_int(rb.sx)); b.nhull = __float_as_int(rb.sy); b.size.z = rb.rbound; }
                            a.c = V3{it.ca[0], it.ca[1], it.ca[2]}; b.c = V3{it.cb[0], it.cb[1], it.cb[2]};
#pragma unroll
                            for (int k = 0; k < 9; k++) { a.R.m[k] = it.Ra[k]; b.R.m[k] = it.Rb[k]; }
                            mpr_begin(m, a, b);
                            active = true;
                            VK_STAT(14, 0);
                        }
                    }
                }
                if (!__any_sync(0xffffffffu, active)) {
                    if (__all_sync(0xffffffffu, exhausted)) break;
                    continue;
                }
#ifdef MOPA_VK_STATS   // [13][k]: trips in phase k; [14]: items begun, active lanes summed over trips, warp trips, kind-pair changes
                if (active) VK_STAT(13, m.phase & 3);
                if ((tid & 31) == 0) { atomicAdd(&g_vk_stats[14][1], (unsigned long long)__popc(__ballot_sync(0xffffffffu, active))); VK_STAT(14, 2); }
                else __ballot_sync(0xffffffffu, active);
#endif
                if (active) {
                    float depth = 0.0f;
                    const int r = mpr_trip<MESH>(m, a, b, &depth);
                    if (r != MPR_RUNNING) {
                        if (r == MPR_PENETRATING && -depth <= thr) { atomicMin(&out[iq], canon); VK_STAT(PC_MPR, 2); }
                        active = false;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) { wcount[4] = 0; wcount[5] = 0; }
        __syncthreads();
    };

    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int q = tile * NQ + tid;
        if (tid < 3) wcount[tid] = 0;
        res[tid] = 0xFFFFFFFFu;
        if (q < n) {
            RowQ rq{qpos + (size_t)q * row_stride};
            fk_state(S, rq, frames, NQ, tid);
        }
        __syncthreads();
        // ---- phase B.  Every lane sweeps (rows beyond n hold stale frames and push nothing) so that the warp can be
        // re-converged explicitly after the divergent push loop.  The pair table is laid out in windows of 32 entries; a run of
        // pairs with the same cull kind and anchor never straddles a window, is padded to a multiple of four and shares the
        // anchor centre.  Each test leaves its verdict in the sign of (cull bound - measured value): exactly the comparison of the
        // filter (a difference of two floats has the sign of the comparison), shifted into the run's mask with one funnel shift.
        // Entries are visited last to first so that entry j of the run ends up in bit j.
        {
            const uint32_t live = q < n ? 0xFFFFFFFFu : 0u;
            uint32_t mw = 0u;                 // survivors of the current window
            // Survivors of one window -> work lists, without per-item atomics: the scene tells which entries of the window
            // belong to which list, every lane counts its survivors per list, one warp scan (three 11-bit counters packed in a
            // word) gives each lane its offsets, three lanes reserve the warp's share of the lists, and the lanes write their
            // items.  Called by all lanes (the run table is uniform).
            auto flush = [&](int base) {      // pairs base .. base + 31 <-> bits of mw
                mw &= live;
                const uint4 wm = S.wmask[base >> 5];
                const uint32_t m0 = mw & wm.x, m1 = mw & wm.y, m2 = mw & wm.z;
                const uint32_t packed = (uint32_t)__popc(m0) | ((uint32_t)__popc(m1) << 11) | ((uint32_t)__popc(m2) << 22);
                if (__any_sync(0xffffffffu, packed != 0u)) {
                    const int lane = tid & 31;
                    uint32_t incl = packed;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31), excl = incl - packed;
                    int reserved = 0;
                    if (lane < WL_COUNT) {
                        const int c = (int)((total >> (11 * lane)) & 0x7FFu);
                        if (c) reserved = atomicAdd(&wcount[lane], c);
                    }
                    const uint32_t item0 = (((uint32_t)tid << 16) | (uint32_t)base) - 1u;   // + __ffs(bits) = (query, pair)
                    auto emit = [&](uint32_t bits, int list, int idx) {
                        uint32_t *dst = wl_base(list) + idx;
                        if (idx + __popc(bits) <= wl_cap(list)) {   // the common case: no bounds check per item
                            while (bits) {
                                VK_STAT(S.pairs[base + __ffs(bits) - 1].cls, 1);
                                *dst++ = item0 + (uint32_t)__ffs(bits);
                                bits &= bits - 1;
                            }
                        } else {
                            const int room = wl_cap(list) - idx;
                            for (int k = 0; bits; k++) {
                                const int p = base + __ffs(bits) - 1;
                                bits &= bits - 1;
                                VK_STAT(S.pairs[p].cls, 1);
                                if (k < room) dst[k] = ((uint32_t)tid << 16) | (uint32_t)p;
                                else eval_item_overflow<MESH>(smem, frames, NQ, tid, p, thr, &res[tid]);
                            }
                        }
                    };
                    emit(m0, WL_CHEAP, __shfl_sync(0xffffffffu, reserved, WL_CHEAP) + (int)(excl & 0x7FFu));
                    emit(m1, WL_BOX, __shfl_sync(0xffffffffu, reserved, WL_BOX) + (int)((excl >> 11) & 0x7FFu));
                    emit(m2, WL_MPR, __shfl_sync(0xffffffffu, reserved, WL_MPR) + (int)(excl >> 22));
                    __syncwarp();
                }
            };
            const int ngroup = S.H->n_group;
#pragma unroll 1
            for (int g = 0; g < ngroup; g++) {
                const CullGroup G = S.groups[g];
                const float *fa = frames + (int)G.anchor_slot * NQ + tid;
                const V3 ac{fa[0], fa[NQ], fa[2 * NQ]};
                const CullEntry *E0 = S.cull + G.first;
                uint32_t culled = 0u;
                if (G.kind == CK_SPHERE_STATIC) {
#pragma unroll 1
                    for (int j = G.count - 4; j >= 0; j -= 4) {
#pragma unroll
                        for (int u = 3; u >= 0; u--) {
                            const CullEntry E = E0[j + u];
                            const V3 d = V3{E.x, E.y, E.z} - ac;
                            culled = __funnelshift_l(__float_as_uint(E.w - dot(d, d)), culled, 1);
                        }
                    }
                } else if (G.kind == CK_SPHERE_MOVING) {
#pragma unroll 1
                    for (int j = G.count - 4; j >= 0; j -= 4) {
#pragma unroll
                        for (int u = 3; u >= 0; u--) {
                            const CullEntry E = E0[j + u];
                            const float *fp = frames + __float_as_int(E.x) * NQ + tid;
                            const V3 d = V3{fp[0], fp[NQ], fp[2 * NQ]} - ac;
                            culled = __funnelshift_l(__float_as_uint(E.w - dot(d, d)), culled, 1);
                        }
                    }
                } else if (G.kind == CK_PLANE) {
#pragma unroll 1
                    for (int j = G.count - 4; j >= 0; j -= 4) {
#pragma unroll
                        for (int u = 3; u >= 0; u--) {
                            const CullEntry E = E0[j + u];
                            culled = __funnelshift_l(__float_as_uint(E.w - dot(V3{E.x, E.y, E.z}, ac)), culled, 1);
                        }
                    }
                }
                const uint32_t cmask = G.count == 32 ? 0xFFFFFFFFu : ((1u << G.count) - 1u);
                mw |= (~culled & cmask) << G.bitpos;
#ifdef MOPA_VK_STATS
                for (int j = 0; j < G.count; j++) VK_STAT(S.pairs[G.first + j].cls, 0);
#endif
                if (G.flush) { flush((int)G.first - (int)G.bitpos); mw = 0u; }
            }
        }
        __syncthreads();
        // ---- phase C: analytic pairs, then box-box
#pragma unroll 1
        for (int list = 0; list < WL_MPR; list++) {
            const int cnt = min(wcount[list], wl_cap(list));
            const uint32_t *items = wl_base(list);
            for (int i = tid; i < cnt; i += NQ) {
                const uint32_t item = items[i];
                const int ql = item >> 16, p = item & 0xFFFF;
                if (!exact && res[ql] != 0xFFFFFFFFu) continue;   // already known to be invalid
                eval_item<MESH, false>(S, frames, NQ, ql, S.pairs[p], thr, &res[ql]);
            }
            if (!exact) __syncthreads();    // the cheap verdicts spare the expensive items of invalid states
        }
        // portal-refinement candidates: conservative pre-test, survivors are queued with their frames
        {
            const int cnt = min(wcount[WL_MPR], wl_cap(WL_MPR));
            const uint32_t *items = wl_base(WL_MPR);
            for (int i = tid; i < cnt; i += NQ) {
                const uint32_t item = items[i];
                const int ql = item >> 16, p = item & 0xFFFF;
                if (!exact && res[ql] != 0xFFFFFFFFu) continue;
                const PairRec pr = S.pairs[p];
                Geom a, b;
                load_geom<MESH>(a, S.recs[pr.ga], frames, NQ, ql);
                load_geom<MESH>(b, S.recs[pr.gb], frames, NQ, ql);
                if (mpr_certainly_separate(a, b)) continue;
                VK_STAT(pr.cls, 3);
                const int slot = atomicAdd(&wcount[4], 1);
                if (slot < mq_cap) {
                    MprItem &it = mq[slot];
                    it.q = tile * NQ + ql; it.p = p;
                    it.ca[0] = a.c.x; it.ca[1] = a.c.y; it.ca[2] = a.c.z; it.cb[0] = b.c.x; it.cb[1] = b.c.y; it.cb[2] = b.c.z;
#pragma unroll
                    for (int k = 0; k < 9; k++) { it.Ra[k] = a.R.m[k]; it.Rb[k] = b.R.m[k]; }
                } else
                    eval_item_overflow<MESH>(smem, frames, NQ, ql, p, thr, &res[ql]);
            }
        }
        __syncthreads();
        if (q < n) out[q] = res[tid];          // raw verdict; queued refinements may still lower it
        __syncthreads();
        if (wcount[4] >= mq_run || tile + (int)gridDim.x >= ntile) run_queue();   // enough work queued, or this was the CTA's last tile
    }
    for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const int q = tile * NQ + tid;
        if (q < n) {
            const uint32_t r = out[q];
            out[q] = (r == 0xFFFFFFFFu) ? 1u : (exact ? ((r + 1u) << 8) : 0u);
        }
    }
}

// Queries per CTA.  One large CTA per SM is preferred over several small ones: its warps walk the phases together, so the code
// an SM executes at any moment is one phase's loop (the kernel is ~130 KB of SASS, the instruction cache next to the SM 32 KB;
// three independent 128-thread CTAs kept three different phases in flight and the kernel waited for instructions), and the
// scene tables exist once per SM.  The frame store (frame floats x 4 B per query) decides which size fits.
static const int VK_NQ_CHOICES[] = {448, 384, 256, 128};

// shared-memory plan for one scene: [blob][frames NQ x ffs][res NQ][counters][work lists]
struct VkPlan { int nq, ffs, cap_cheap, cap_box, cap_mpr, mq_cap, mq_run; size_t smem; int per_sm; };
static VkPlan validity_plan(const SceneHeader &H, long long n = -1, int sm_count = 148) {
    VkPlan P;
    P.ffs = H.frame_floats;
    const size_t sm_total = 228 * 1024, reserve = 1024;
    size_t lists = 0, fixed = 0;
    P.nq = 128; P.per_sm = 0;
    for (int nq : VK_NQ_CHOICES) {
        fixed = (size_t)H.blob_bytes + (size_t)P.ffs * nq * 4 + (size_t)nq * 4 + 192 + 16;
        if (nq > 128) {   // one CTA per SM, lists of >= 4 KB per 128 queries; small batches keep the small CTAs (more SMs busy)
            if (n >= 0 && n < (long long)nq * sm_count) continue;
            const size_t budget = 227 * 1024 - reserve, need = 4096 * (size_t)(nq / 128);
            if (budget >= fixed + need) { P.nq = nq; P.per_sm = 1; lists = budget - fixed; break; }
            continue;
        }
        for (int per_sm = 4; per_sm >= 1; per_sm--) {   // most resident CTAs that still leave >= 4 KB of work lists each
            const size_t budget = sm_total / per_sm - reserve;
            if (budget > 227 * 1024) continue;
            if (budget >= fixed + 4096) { P.per_sm = per_sm; lists = budget - fixed; break; }
        }
    }
    if (P.per_sm < 1) { P.per_sm = 1; lists = 4096; }
    const size_t lists_max = 16384 * (size_t)(P.nq / 128);
    if (lists > lists_max) lists = lists_max;
    const int items = (int)(lists / 4) & ~3;
    P.cap_box = items / 6; P.cap_mpr = items / 2; P.cap_cheap = items - P.cap_box - P.cap_mpr;   // ~ 4 : 1.4 : 5.6 items per query
    P.mq_run = 8 * P.nq;    // queued refinements per CTA before phase D runs (8 per thread)
    P.mq_cap = P.mq_run + P.cap_mpr;
    P.smem = fixed + (size_t)items * 4;
    return P;
}
size_t validity_smem_bytes(const SceneHeader &H) { return validity_plan(H).smem; }

// per (device, stream) slab of queued refinement items: launches on one stream are ordered, launches on different streams
// (the rollout's main and planner streams) must not share it
static MprItem *validity_scratch(cudaStream_t stream, size_t bytes, cudaError_t &err) {
    struct Slab { void *p; size_t bytes; };
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, Slab> slabs;
    int dev = 0;
    err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    Slab &s = slabs[std::make_pair(dev, stream)];
    if (s.bytes < bytes) {
        if (s.p) { err = cudaStreamSynchronize(stream); if (err != cudaSuccess) return nullptr; cudaFree(s.p); s.p = nullptr; s.bytes = 0; }
        err = cudaMalloc(&s.p, bytes);
        if (err != cudaSuccess) { s.p = nullptr; return nullptr; }
        s.bytes = bytes;
    }
    return (MprItem *)s.p;
}

template <int NQ, bool MESH>
static cudaError_t launch_is_valid_t(const VkPlan &P, const unsigned char *d_blob, const SceneHeader &H, const float *d_qpos, int row_stride, int n,
                                     uint32_t *d_out, int exact, int sm_count, cudaStream_t stream, const int *d_n, int d_n_mult) {
    static bool attr_set = false;
    auto kern = is_valid_kernel<NQ, MESH>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int ntile = (n + NQ - 1) / NQ;
    int grid = sm_count * P.per_sm;
    if (grid > ntile) grid = ntile;
    cudaError_t e = cudaSuccess;
    MprItem *mq = validity_scratch(stream, (size_t)sm_count * P.per_sm * P.mq_cap * sizeof(MprItem), e);
    if (!mq) return e;
    kern<<<grid, NQ, P.smem, stream>>>(d_blob, H.blob_bytes, d_qpos, row_stride, n, d_out, exact, d_n, d_n_mult, P.ffs, P.cap_cheap, P.cap_box, P.cap_mpr,
                                       mq, P.mq_cap, P.mq_run);
    return cudaGetLastError();
}

cudaError_t launch_is_valid(const unsigned char *d_blob, const SceneHeader &H, const float *d_qpos, int row_stride, int n,
                            uint32_t *d_out, int exact, int sm_count, cudaStream_t stream, const int *d_n, int d_n_mult) {
    if (n <= 0) return cudaSuccess;
    const VkPlan P = validity_plan(H, n, sm_count);
    const bool mesh = H.n_hull_vert > 0;   // scenes with mesh colliders run the instantiation that carries the hull support
#define VK_LAUNCH(NQ_) (mesh ? launch_is_valid_t<NQ_, true>(P, d_blob, H, d_qpos, row_stride, n, d_out, exact, sm_count, stream, d_n, d_n_mult) \
                             : launch_is_valid_t<NQ_, false>(P, d_blob, H, d_qpos, row_stride, n, d_out, exact, sm_count, stream, d_n, d_n_mult))
    if (P.nq == 448) return VK_LAUNCH(448);
    if (P.nq == 384) return VK_LAUNCH(384);
    if (P.nq == 256) return VK_LAUNCH(256);
    return VK_LAUNCH(128);
#undef VK_LAUNCH
}

#ifdef MOPA_VK_STATS
extern "C" int mopa_debug_vk_stats(unsigned long long *out64) { return (int)cudaMemcpyFromSymbol(out64, g_vk_stats, sizeof(g_vk_stats)); }
#endif

}  // namespace mopa
