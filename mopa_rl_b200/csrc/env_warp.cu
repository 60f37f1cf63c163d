// env.step on sm_100a, ONE WARP PER ENVIRONMENT (the production path).
//
// Same mathematics as dyn.cuh / contact.cuh (one mj_step equivalent per substep, 75 substeps per
// env.step, reward / observation / _after_step fused; reference call sites in env_kernel.cu), but
// every stage is spread over the 32 lanes of the warp that owns the environment and the
// per-environment matrices live in a shared-memory workspace instead of per-thread local memory:
//
//   kinematics chain            uniform (all lanes), frames -> shared
//   body inertias, RNE terms    lane = body          chain accumulations: lanes = vector components
//   CRBA rows, bias, forces     lane = dof
//   Cholesky (M, M + hD)        lane = row, right-looking
//   contact broad phase         lanes stride the candidate pairs, ballot compaction
//   contact narrow phase        lane = surviving pair
//   constraint rows             lane = row: Jacobian, Y = L^-1 J^T (forward substitution), A = Y Y^T
//   projected Gauss-Seidel      lane r owns residual g_r; row updates are broadcast by shuffle
//
// 4096 environments = 4096 warps, i.e. 8-9 resident warps per SM instead of < 1 for the
// thread-per-env kernel.  Reductions change summation order, so agreement with the CPU oracle is
// at rounding level (tests: 1e-5 absolute on qpos / qvel as BASELINE.json states).
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

#include "../../include/mopa_b200.h"
#include "dyn.cuh"
#include "contact.cuh"
#include "env_state.h"

namespace mopa {

constexpr int WD = DMAXD;       // dofs
// WC = constraint-row capacity (one lane each, <= 32), WC / 3 = contact points kept per substep: template parameters
constexpr int WCAND = 48;       // broad-phase survivors kept per substep
constexpr int YS = WD + 1;      // padded row stride of Y (bank-conflict free, lane = row)
constexpr int NTRI = WD * (WD + 1) / 2;   // lower-triangular storage of M and its Cholesky factors
#define TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

// WB / WG: body and contact-geom capacity of the workspace (two instantiations: small scenes such as
// Sawyer-Push fit 14 warps per SM, which turns 4096 envs into exactly two waves on 148 SMs)
template <int WB>
struct WarpKin {
    double xpos[WB][3], xquat[WB][4], xmat[WB][9];
    double S[WD][6];
    double vel[WB][6], frc[WB][6];
    double inert[WB][13];
};
template <int WB, int WG, int WC>
struct WarpWS {
    static constexpr int WCP = WC / 3;
    double q[36], v[36];
    union {              // the kinematic arrays are dead once the constraint Jacobians exist: A reuses them
        WarpKin<WB> k;
        struct {
            double A[WC * YS];   // solver scratch, row r at A + r * YS (padded stride: lane = row accesses are bank-conflict free): y = L^-1 J_r^T, then K_r (Hessian assembly)
            // Behind A lie k.vel / k.frc / k.inert, dead once the inertia matrix exists: the factor of M + h * diag(damping)
            // (implicit joint damping of the Euler update) is produced there by the upper half-warp while the lower one
            // factors M, and waits for the velocity update at the end of the substep.
            double L2[NTRI], invd2[WD];
            alignas(16) double col2[4 * (WD + 2)];   // column buffers of w_factor_solve: per (half-warp, diagonal block)
        };
    };
    static_assert(WC * YS >= WB * 16 + WD * 6, "A must cover xpos / xquat / xmat / S so that L2 only overlaps arrays dead after stage 3");
    double kxpos[4][3], kxquat[2][4], kxmat[4][9];   // frames of the bodies the env epilogue reads (mjData after mj_step)
    double M[NTRI], L[NTRI], invd[WD];
    double qd[WD], bias[WD], tau[WD], qacc0[WD], a[WD], rhs[WD], bias_prev[WD], ctrl[DMAXA], z[WD];
    double Y[WC * YS];
    alignas(16) double f[WC];   // gradient of the penalties per row; also the column buffer of w_factor_solve
    double gpos[WG][3];
    static constexpr int WGM = WG <= 32 ? 2 : DMAXGM;
    double gmat[WGM][9];     // world rotation of the moving geoms whose local rotation is not the identity
    double cpos[WCP][3], cn[WCP][3], cdist[WCP], cmargin[WCP], cmu[WCP];
    int cga[WCP], cgb[WCP], csig[WCP];
    double wa[WD];           // warm start: acceleration of the previous substep (mjData.qacc_warmstart)
    int wn;                  // ... and whether it exists
    int blk[WD];             // first dof of the kinematic tree that owns each dof
    unsigned short cand[WCAND];
    int ncand, ncp;
    int blk0;                // dofs [0, blk0) and [blk0, nd) are two kinematic trees (0: any other structure): M is block diagonal
};

#define FULL 0xffffffffu

// scene constants: uniform reads go through the constant cache instead of global loads
constexpr int ENV_MODEL_SLOTS = 2;
__constant__ DynDev c_models[ENV_MODEL_SLOTS];
// Task tables live in the constant bank too: a kernel parameter struct that is indexed with a lane id gets copied to
// local memory by every thread (1.2 KB x 131 k threads = 150 MB of writes per launch in round 1's ncu capture).
__constant__ mopa_sawyer_task c_tasks[ENV_MODEL_SLOTS];
struct EnvTune { int prof, sync_mask; };
__constant__ EnvTune c_tune;
__device__ unsigned long long g_prof[32];   // [2k] work before stage barrier k, [2k+1] wait at it; 20 = PGS sweeps, 21 = rows, 22 = substeps

__device__ __forceinline__ double shfl_d(double x, int src) { return __shfl_sync(FULL, x, src); }
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}

// ---- Cholesky factorisation + solve, register resident.  Lane i keeps row i of the lower triangle in
// registers (template recursion = full unrolling: every row[] index is a compile-time constant).
// Column j: lane j publishes d_jj and its partially reduced right-hand side through a small shared
// buffer; every lane scales by rsqrt(d_jj) (no fp64 divide / sqrt on the critical path), publishes
// its l_ij, and the trailing update reads the column back with broadcast 128-bit loads.  The forward
// substitution L y = b rides along with the elimination.  The factor is then written to shared memory
// (packed lower triangle, plus 1 / L_ii) for the backward substitution and for later triangular
// solves of the caller.  Rows / columns nd..WD-1 are identity padding (no bound checks inside).
//   (M + hs * diag) = L L^T,  x <- (M + hs * diag)^-1 x      x: shared, nd entries; col: shared, WD + 2 doubles, 16-byte aligned
template <int N, int J, int K>
__device__ __forceinline__ void wfs_update(double (&row)[N], const double (&c)[N], int lj) {
    if constexpr (K < N) {
        if (lj >= K) row[K] -= row[J] * c[K];
        wfs_update<N, J, K + 1>(row, c, lj);
    }
}
template <int N, int K2>
__device__ __forceinline__ void wfs_fetch(double (&c)[N], const double *col) {
    if constexpr (2 * K2 < N) {
        const double2 v = reinterpret_cast<const double2 *>(col)[K2];
        c[2 * K2] = v.x;
        if constexpr (2 * K2 + 1 < N) c[2 * K2 + 1] = v.y;
        wfs_fetch<N, K2 + 1>(c, col);
    }
}
template <int N, int J>
__device__ __forceinline__ void wfs_column(double (&row)[N], double &xr, double &rme, int lj, bool act, double *col) {
    if constexpr (J < N) {
        if (lj == J && act) { col[WD] = row[J]; col[WD + 1] = xr; }
        __syncwarp();
        const double djj = col[WD], rinv = rsqrt(djj), yj = col[WD + 1] * rinv;
        if (lj == J) { row[J] = djj * rinv; rme = rinv; xr = yj; }
        else if (lj > J) { row[J] *= rinv; xr -= row[J] * yj; if (act) col[lj] = row[J]; }
        __syncwarp();
        if constexpr (J + 1 < N) {
            double c[N];
            wfs_fetch<N, (J + 1) / 2>(c, col);
            wfs_update<N, J, J + 1>(row, c, lj);
        }
        wfs_column<N, J + 1>(row, xr, rme, lj, act, col);
    }
}
template <int N, int K>
__device__ __forceinline__ void wfs_load(double (&row)[N], const double *Msrc, double dd, int li, int s, bool live) {
    if constexpr (K < N) {
        row[K] = (live && s + K <= li) ? Msrc[TRI(li, s + K)] : 0.0;
        if (s + K == li) row[K] = live ? row[K] + dd : 1.0;
        wfs_load<N, K + 1>(row, Msrc, dd, li, s, live);
    }
}
template <int N, int K>
__device__ __forceinline__ void wfs_store(const double (&row)[N], double *Lout, int li, int s, bool live) {
    if constexpr (K < N) {
        if (live && s + K <= li) Lout[TRI(li, s + K)] = row[K];
        wfs_store<N, K + 1>(row, Lout, li, s, live);
    }
}
// Lanes 0..15 factor Msrc + hs * diag and solve for x.  When Lout2 is given, lanes 16..31 factor Msrc + hs2 * diag2 in the same
// instruction stream (factor only: Lout2 / invd2 feed w_solve_stored later).
// BLK0 > 0: the matrix is block diagonal with blocks [0, BLK0) and [BLK0, nd) - two kinematic trees, e.g. the arm and the free
// object, and no constraint row that couples them - and the two blocks are eliminated side by side: a lane keeps the part of its
// row that lies inside its block (register k <-> column block start + k), every block has its own column buffer, and the sweep
// runs over max(block size) columns instead of WD (9 instead of 16 on the Sawyer scenes, with 9 row registers instead of 16).
// The entries outside the blocks are exact zeros in the dense elimination too (x - 0 * y = x), so the factor is the same bit for
// bit; they are neither written nor read here.  `col`: (WD + 2) doubles per (half-warp, block) in use, 16-byte aligned.
template <int BLK0>
__device__ __noinline__ void w_factor_solve(const double *Msrc, const double *diag, double hs, int nd, double *Lout, double *invd, double *x,
                                            double *col, int lane, double *Lout2 = nullptr, double *invd2 = nullptr,
                                            const double *diag2 = nullptr, double hs2 = 0.0) {
    constexpr int N = BLK0 > 0 ? (BLK0 > WD - BLK0 ? BLK0 : WD - BLK0) : WD;   // columns of the largest block
    double row[N];
    const int li = lane & 15, half = lane >> 4;
    const int b = (BLK0 > 0 && li >= BLK0) ? 1 : 0, s = b ? BLK0 : 0, lj = li - s;
    const bool act = half == 0 || Lout2 != nullptr;   // this half-warp owns a matrix
    const bool live = act && li < nd;
    const double *dg = half ? diag2 : diag;
    wfs_load<N, 0>(row, Msrc, (dg && live) ? (half ? hs2 : hs) * dg[li] : 0.0, li, s, live);
    double xr = (live && half == 0) ? x[li] : 0.0, rme = 0.0;
    wfs_column<N, 0>(row, xr, rme, lj, act, col + (act ? (half * 2 + b) * (WD + 2) : 0));
    wfs_store<N, 0>(row, half ? Lout2 : Lout, li, s, live);
    if (live) (half ? invd2 : invd)[li] = rme;
    __syncwarp();
    // backward substitution L^T x = y: lane i broadcasts x_i, the lanes below it (in its block) eliminate
    for (int i = nd - 1; i >= 0; i--) {
        const double xi = shfl_d(xr * rme, i);
        const int si = (BLK0 > 0 && i >= BLK0) ? BLK0 : 0;
        if (lane == i) xr = xi;
        else if (lane < i && lane >= si) xr -= Lout[TRI(i, lane)] * xi;
    }
    if (lane < nd) x[lane] = xr;
    __syncwarp();
}
constexpr int SAWYER_BLK0 = 9;   // the arm + gripper tree of the Sawyer scenes has 9 dofs; the free object follows
constexpr int SMALL_ND = 8;      // scenes with at most 8 dofs (PusherObstacle-v0): one block of 8 columns, the second block is empty - half the sweep
// x <- (L L^T)^-1 x with a factor w_factor_solve stored earlier (same operation order as its fused forward substitution)
__device__ __noinline__ void w_solve_stored(const double *L, const double *invd, int nd, double *x, int lane, int blk0) {
    const bool live = lane < nd;
    double xr = live ? x[lane] : 0.0;
    const double rme = live ? invd[lane] : 0.0;
    const int sl = (blk0 > 0 && lane >= blk0) ? blk0 : 0;   // first column of this lane's block
    for (int i = 0; i < nd; i++) {
        const double yi = shfl_d(xr * rme, i);
        if (lane == i) xr = yi;
        else if (lane > i && live && i >= sl) xr -= L[TRI(lane, i)] * yi;
    }
    for (int i = nd - 1; i >= 0; i--) {
        const double xi = shfl_d(xr * rme, i);
        const int si = (blk0 > 0 && i >= blk0) ? blk0 : 0;
        if (lane == i) xr = xi;
        else if (lane < i && lane >= si) xr -= L[TRI(i, lane)] * xi;
    }
    if (live) x[lane] = xr;
    __syncwarp();
}

// ---- box-box face contact, warp cooperative: same arithmetic per vertex as box_box_face (contact.cuh), with the
// polygon distributed over the lanes (lane q = vertex q, at most 8).  Each Sutherland-Hodgman stage: every lane tests
// its vertex and its successor's, emits 0-2 vertices, and the survivors are compacted in order through `tmp` (shared,
// >= 8 x 3 doubles).  The <= 4 deepest points are written to out[] in vertex order; returns their number (all lanes).
__device__ __noinline__ int box_box_face_warp(CPoint *out, const CGeom &g1, const CGeom &g2, int code, double margin, double (*tmp)[3], int lane) {
    const CGeom &gr = code < 3 ? g1 : g2, &gi = code < 3 ? g2 : g1;
    const int ax = code < 3 ? code : code - 3;
    double n[3], dd[3] = {gi.c[0] - gr.c[0], gi.c[1] - gr.c[1], gi.c[2] - gr.c[2]};
    c_colk(n, gr.R, ax);
    if (d_dot(n, dd) < 0) for (int k = 0; k < 3; k++) n[k] = -n[k];
    int iax = 0;
    double mind = 1e30, isg = 1;
    for (int k = 0; k < 3; k++) {
        double a[3];
        c_colk(a, gi.R, k);
        const double dn = d_dot(a, n);
        if (-fabs(dn) < mind) { mind = -fabs(dn); iax = k; isg = dn > 0 ? -1.0 : 1.0; }
    }
    const int u = (iax + 1) % 3, v = (iax + 2) % 3;
    double P[3] = {0, 0, 0};
    int np = 4;
    if (lane < 4) {
        double fa[3], ua[3], va[3];
        c_colk(fa, gi.R, iax); c_colk(ua, gi.R, u); c_colk(va, gi.R, v);
        const double su = (lane == 0 || lane == 3) ? 1.0 : -1.0, sv = lane < 2 ? 1.0 : -1.0;
        for (int k = 0; k < 3; k++) P[k] = gi.c[k] + isg * gi.size[iax] * fa[k] + su * gi.size[u] * ua[k] + sv * gi.size[v] * va[k];
    }
    const int ru = (ax + 1) % 3, rv = (ax + 2) % 3;
    for (int side = 0; side < 4 && np > 0; side++) {
        double pa[3];
        c_colk(pa, gr.R, side < 2 ? ru : rv);
        const double sg = (side % 2) ? -1.0 : 1.0, lim = gr.size[side < 2 ? ru : rv];
        const bool live = lane < np;
        const double dp = sg * ((P[0] - gr.c[0]) * pa[0] + (P[1] - gr.c[1]) * pa[1] + (P[2] - gr.c[2]) * pa[2]) - lim;
        const int nxt = live ? (lane + 1 == np ? 0 : lane + 1) : lane;
        const double dq = shfl_d(dp, nxt), Q0 = shfl_d(P[0], nxt), Q1 = shfl_d(P[1], nxt), Q2 = shfl_d(P[2], nxt);
        const bool keepP = live && dp <= 0, cross = live && ((dp <= 0) != (dq <= 0));
        const int cnt = (keepP ? 1 : 0) + (cross ? 1 : 0);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
        const int off = incl - cnt;
        if (keepP) { tmp[off][0] = P[0]; tmp[off][1] = P[1]; tmp[off][2] = P[2]; }
        if (cross) {
            const double t = dp / (dp - dq);
            const int o2 = off + (keepP ? 1 : 0);
            tmp[o2][0] = P[0] + t * (Q0 - P[0]); tmp[o2][1] = P[1] + t * (Q1 - P[1]); tmp[o2][2] = P[2] + t * (Q2 - P[2]);
        }
        np = __shfl_sync(FULL, incl, 7);
        __syncwarp();
        if (lane < np) { P[0] = tmp[lane][0]; P[1] = tmp[lane][1]; P[2] = tmp[lane][2]; }
        __syncwarp();
    }
    const double depth = (P[0] - gr.c[0]) * n[0] + (P[1] - gr.c[1]) * n[1] + (P[2] - gr.c[2]) * n[2] - gr.size[ax];
    bool keep = lane < np && depth < margin;
    unsigned km = __ballot_sync(FULL, keep);
    while (__popc(km) > 4) {   // drop the shallowest point (first one on ties), like the serial selection
        double bv = keep ? depth : -1e300;
        int bi = lane;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(FULL, bv, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        bi = __shfl_sync(FULL, bi, 0);
        if (lane == bi) keep = false;
        km = __ballot_sync(FULL, keep);
    }
    const double flip = (&gr == &g1) ? 1.0 : -1.0;
    if (keep) {
        const int slot = __popc(km & ((1u << lane) - 1));
        out[slot].dist = depth;
        for (int k = 0; k < 3; k++) { out[slot].n[k] = flip * n[k]; out[slot].pos[k] = P[k] - n[k] * 0.5 * depth; }
    }
    __syncwarp();
    return __popc(km);
}

// Out-of-line wrappers keep one copy of the big routines in the instruction stream (the kernel is
// instruction-cache bound otherwise: 9 warps per SM walk different phases of a long program).
__device__ __noinline__ int pair_contacts_ni(CPoint *out, const CGeom &ga, const CGeom &gb, double margin) {
    return pair_contacts(out, ga, gb, margin);
}
__device__ __noinline__ void kbi_ni(const DynDev &m, const double *solref, const double *solimp, double pos, double margin, double &K,
                                    double &B, double &imp) {
    d_kbi(m, solref, solimp, pos, margin, K, B, imp);
}

// mj_integratePos with the velocities in W.qd: hinge / slide joints linear, free joints by the exponential map of the angular velocity
template <int WB, int WG, int WC>
__device__ __forceinline__ void w_integrate_pos(const DynDev &m, WarpWS<WB, WG, WC> &W, double h, int lane) {
    if (lane < m.nb) {
        const int i = lane, jt = m.b_jtype[i], da = m.b_dadr[i], a = m.b_qadr[i];
        if (jt == 2 || jt == 3) W.q[a] += h * W.qd[da];
        else if (jt == 0) {
            for (int k = 0; k < 3; k++) W.q[a + k] += h * W.qd[da + k];
            const double w[3] = {W.qd[da + 3], W.qd[da + 4], W.qd[da + 5]}, n = sqrt(d_dot(w, w)), ang = n * h;
            if (ang > 0) {
                const double sn = sin(0.5 * ang) / n, dq[4] = {cos(0.5 * ang), w[0] * sn, w[1] * sn, w[2] * sn};
                double qn[4], q0[4] = {W.q[a + 3], W.q[a + 4], W.q[a + 5], W.q[a + 6]};
                d_qmul(qn, q0, dq);
                const double nn = sqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
                for (int k = 0; k < 4; k++) W.q[a + 3 + k] = qn[k] / nn;
            }
        }
    }
    __syncwarp();
}

// One mj_step for the warp's environment.  W.q / W.v hold the state, W.ctrl the controls,
// `comp` the dofs whose qfrc_applied is the previous qfrc_bias.  smode: 0 kinematics + bias only (sim.forward), 1 one mj_step
// with the semi-implicit Euler integrator, 2 forward dynamics only: W.rhs <- qacc = M^-1 (tau + J^T f) (a stage of mj_RungeKutta),
// 3 kinematics + contact detection only (ncon of a candidate reset state).
// `sync`: the warps of a CTA walk the stages in step (CTA barrier between stages) so that they share
// instruction-cache lines; `active` = false warps only take part in the barriers.
// profiling / tuning hooks (MOPA_ENV_PROF=1, MOPA_ENV_SYNC_MASK=bits): per-stage clock64 sums, lane 0 of every warp
#define PROF_MARK(id) do { if (c_tune.prof) { const long long t_ = clock64(); if (lane == 0) atomicAdd(&g_prof[id], (unsigned long long)(t_ - t_last)); t_last = t_; } } while (0)
#define STAGE_SYNC(k) do { PROF_MARK(2 * (k)); if (sync && ((c_tune.sync_mask >> (k)) & 1)) __syncthreads(); PROF_MARK(2 * (k) + 1); } while (0)
// One mj_step is three out-of-line stages (each gets its own register allocation: as one function the live ranges of
// the kinematics, the contact generator and the solver overlapped and spilled): dynamics terms up to the unconstrained
// acceleration, constraint discovery (joint limits, contact points), constraint solve + integration.  State passes through
// the warp's shared-memory workspace; the barrier-only path of inactive warps lives in the driver w_substep.
// ND: the scene's dof count as a compile-time constant (0 = read it from the model).  All three Sawyer scenes have 15 simulated dofs;
// with the count fixed the per-dof / per-row loops unroll and their index arithmetic folds (the kernel is bound by the dependent
// instruction chain of one env.step, and 2/3 of its instructions were loop control and addressing).
template <int WB, int WG, int WC, int ND, int NB>   // NB: likewise the number of simulated bodies (14 on the push scene)
__device__ __noinline__ bool w_stage_dynamics(const DynDev &m, const DynDev *__restrict__ mg, WarpWS<WB, WG, WC> &W, unsigned comp, int smode, int lane, const int4 keep_bodies, bool sync) {
    const int nb = NB > 0 ? NB : m.nb, nd = ND > 0 ? ND : m.nd;
    long long t_last = c_tune.prof ? clock64() : 0;
    if (lane < nd) W.qd[lane] = W.v[m.d_vadr[lane]];
    __syncwarp();
    // ---- kinematics, lane = body, in registers: (1) local transform of the body incl. its joint (the
    // expensive sin / cos run in parallel), (2) world frames by pointer jumping over the parent links
    // (log2(depth) rounds of shuffle + compose instead of a serial chain), (3) joint motion axes.
    {
        const int i = lane < nb ? lane : 0;
        const int jt = lane < nb ? m.b_jtype[i] : -1;
        int anc = -1;
        double lp[3] = {0, 0, 0}, lq[4] = {1, 0, 0, 0};
        if (lane < nb) {
            for (int k = 0; k < 3; k++) lp[k] = m.b_pos[i][k];
            for (int k = 0; k < 4; k++) lq[k] = m.b_quat[i][k];
            if (jt == 3) {
                const double ja[3] = {m.b_jaxis[i][0], m.b_jaxis[i][1], m.b_jaxis[i][2]}, jp[3] = {m.b_jpos[i][0], m.b_jpos[i][1], m.b_jpos[i][2]};
                const double ang = W.q[m.b_qadr[i]] - m.b_qpos0[i], sn = sin(0.5 * ang), cs = cos(0.5 * ang);
                const double ql[4] = {cs, sn * ja[0], sn * ja[1], sn * ja[2]};
                double qn[4];
                d_qmul(qn, lq, ql);
                if (jp[0] != 0.0 || jp[1] != 0.0 || jp[2] != 0.0) {   // rotation about an offset anchor
                    double R0[9], R1[9], t0[3], t1[3];
                    d_q2m(R0, lq); d_q2m(R1, qn);
                    d_mv(t0, R0, jp); d_mv(t1, R1, jp);
                    for (int k = 0; k < 3; k++) lp[k] += t0[k] - t1[k];
                }
                for (int k = 0; k < 4; k++) lq[k] = qn[k];
            } else if (jt == 2) {
                const double ja[3] = {m.b_jaxis[i][0], m.b_jaxis[i][1], m.b_jaxis[i][2]};
                double R0[9], ax[3];
                d_q2m(R0, lq);
                d_mv(ax, R0, ja);
                const double dq = W.q[m.b_qadr[i]] - m.b_qpos0[i];
                for (int k = 0; k < 3; k++) lp[k] += ax[k] * dq;
            } else if (jt == 0) {   // free joint: absolute pose
                const int a = m.b_qadr[i];
                for (int k = 0; k < 3; k++) lp[k] = W.q[a + k];
                const double n = sqrt(W.q[a + 3] * W.q[a + 3] + W.q[a + 4] * W.q[a + 4] + W.q[a + 5] * W.q[a + 5] + W.q[a + 6] * W.q[a + 6]);
                for (int k = 0; k < 4; k++) lq[k] = W.q[a + 3 + k] / n;
            }
            if (jt != 0) {
                anc = m.b_parent[i];
                if (anc < 0) {   // tree root: fold the static parent's world frame into the local transform
                    const double Pq[4] = {m.b_rootquat[i][0], m.b_rootquat[i][1], m.b_rootquat[i][2], m.b_rootquat[i][3]};
                    double PM[9], t[3], qn[4];
                    d_q2m(PM, Pq);
                    d_mv(t, PM, lp);
                    for (int k = 0; k < 3; k++) lp[k] = m.b_rootpos[i][k] + t[k];
                    d_qmul(qn, Pq, lq);
                    for (int k = 0; k < 4; k++) lq[k] = qn[k];
                }
            }
        }
        for (int round = 0; round < 5; round++) {
            const bool has = anc >= 0;
            if (!__any_sync(FULL, has)) break;
            const int src = has ? anc : lane;
            double ap[3], aq[4];
#pragma unroll
            for (int k = 0; k < 3; k++) ap[k] = shfl_d(lp[k], src);
#pragma unroll
            for (int k = 0; k < 4; k++) aq[k] = shfl_d(lq[k], src);
            const int aanc = __shfl_sync(FULL, anc, src);
            if (has) {   // T <- T_anc o T
                double PM[9], t[3], qn[4];
                d_q2m(PM, aq);
                d_mv(t, PM, lp);
                for (int k = 0; k < 3; k++) lp[k] = ap[k] + t[k];
                d_qmul(qn, aq, lq);
                for (int k = 0; k < 4; k++) lq[k] = qn[k];
                anc = aanc;
            }
        }
        if (lane < nb) {
            double R[9];
            d_q2m(R, lq);
#pragma unroll
            for (int k = 0; k < 3; k++) W.k.xpos[i][k] = lp[k];
#pragma unroll
            for (int k = 0; k < 4; k++) W.k.xquat[i][k] = lq[k];
#pragma unroll
            for (int k = 0; k < 9; k++) W.k.xmat[i][k] = R[k];
            const int da = m.b_dadr[i];
            const double *pos = lp;
            if (jt == 3) {
                const double ja[3] = {m.b_jaxis[i][0], m.b_jaxis[i][1], m.b_jaxis[i][2]}, jp[3] = {m.b_jpos[i][0], m.b_jpos[i][1], m.b_jpos[i][2]};
                double ax[3], t[3], anchor[3], cr[3];
                d_mv(ax, R, ja);
                d_mv(t, R, jp);
                for (int k = 0; k < 3; k++) anchor[k] = pos[k] + t[k];
                d_cross(cr, anchor, ax);
                for (int k = 0; k < 3; k++) { W.k.S[da][k] = ax[k]; W.k.S[da][3 + k] = cr[k]; }
            } else if (jt == 2) {
                const double ja[3] = {m.b_jaxis[i][0], m.b_jaxis[i][1], m.b_jaxis[i][2]};
                double ax[3];
                d_mv(ax, R, ja);
                for (int k = 0; k < 3; k++) { W.k.S[da][k] = 0; W.k.S[da][3 + k] = ax[k]; }
            } else if (jt == 0) {
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    for (int c = 0; c < 6; c++) W.k.S[da + a][c] = 0;
                    W.k.S[da + a][3 + a] = 1;
                    const double e[3] = {R[a], R[3 + a], R[6 + a]};
                    double cr[3];
                    d_cross(cr, pos, e);
                    for (int c = 0; c < 3; c++) { W.k.S[da + 3 + a][c] = e[c]; W.k.S[da + 3 + a][3 + c] = cr[c]; }
                }
            }
        }
    }
    __syncwarp();
    if (lane < 4) {
        const int b = lane == 0 ? keep_bodies.x : (lane == 1 ? keep_bodies.y : (lane == 2 ? keep_bodies.z : keep_bodies.w));
        for (int k = 0; k < 3; k++) W.kxpos[lane][k] = W.k.xpos[b][k];
        if (lane < 2) for (int k = 0; k < 4; k++) W.kxquat[lane][k] = W.k.xquat[b][k];
        for (int k = 0; k < 9; k++) W.kxmat[lane][k] = W.k.xmat[b][k];
    }
    STAGE_SYNC(1);   // 1: kinematics done
    // ---- spatial inertia about the origin, lane = body
    if (lane < nb) {
        const int i = lane;
        double R[9], c[3], t[3], Mi[9], Ri[9], Iw[9], ip[3] = {m.b_ipos[i][0], m.b_ipos[i][1], m.b_ipos[i][2]};
        double iq[4] = {m.b_iquat[i][0], m.b_iquat[i][1], m.b_iquat[i][2], m.b_iquat[i][3]};
        for (int k = 0; k < 9; k++) R[k] = W.k.xmat[i][k];
        d_mv(t, R, ip);
        for (int k = 0; k < 3; k++) c[k] = W.k.xpos[i][k] + t[k];
        d_q2m(Mi, iq);
        for (int r = 0; r < 3; r++)
            for (int cc = 0; cc < 3; cc++) Ri[3 * r + cc] = R[3 * r] * Mi[cc] + R[3 * r + 1] * Mi[3 + cc] + R[3 * r + 2] * Mi[6 + cc];
        const double i0 = m.b_inertia[i][0], i1 = m.b_inertia[i][1], i2 = m.b_inertia[i][2];
        for (int r = 0; r < 3; r++)
            for (int cc = 0; cc < 3; cc++) Iw[3 * r + cc] = Ri[3 * r] * i0 * Ri[3 * cc] + Ri[3 * r + 1] * i1 * Ri[3 * cc + 1] + Ri[3 * r + 2] * i2 * Ri[3 * cc + 2];
        const double ms = m.b_mass[i], cc2 = d_dot(c, c);
        W.k.inert[i][0] = ms;
        for (int k = 0; k < 3; k++) W.k.inert[i][1 + k] = ms * c[k];
        for (int r = 0; r < 3; r++)
            for (int cc = 0; cc < 3; cc++) W.k.inert[i][4 + 3 * r + cc] = Iw[3 * r + cc] + ms * ((r == cc ? cc2 : 0) - c[r] * c[cc]);
    }
    __syncwarp();
    // ---- velocities down the chain, lanes = components
    for (int i = 0; i < nb; i++) {
        const int p = m.b_parent[i], jt = m.b_jtype[i], da = m.b_dadr[i];
        if (lane < 6) {
            double x = p >= 0 ? W.k.vel[p][lane] : 0.0;
            const int ndj = jt < 0 ? 0 : (jt == 0 ? 6 : 1);
            for (int k = 0; k < ndj; k++) x += W.k.S[da + k][lane] * W.qd[da + k];
            W.k.vel[i][lane] = x;
        }
        __syncwarp();
    }
    // ---- RNE: local acceleration terms (lane = body), chain sum, body forces, backward sum, bias
    if (lane < nb) {
        const int i = lane, jt = m.b_jtype[i], da = m.b_dadr[i];
        Sv6 vi, acc;
        for (int k = 0; k < 3; k++) { vi.w[k] = W.k.vel[i][k]; vi.v[k] = W.k.vel[i][3 + k]; acc.w[k] = 0; acc.v[k] = 0; }
        const int ndj = jt < 0 ? 0 : (jt == 0 ? 6 : 1);
        for (int k = 0; k < ndj; k++) {
            if (jt == 0 && k < 3) continue;
            Sv6 s, sd;
            for (int c = 0; c < 3; c++) { s.w[c] = W.k.S[da + k][c]; s.v[c] = W.k.S[da + k][3 + c]; }
            sv_cross_motion(sd, vi, s);
            for (int c = 0; c < 3; c++) { acc.w[c] += sd.w[c] * W.qd[da + k]; acc.v[c] += sd.v[c] * W.qd[da + k]; }
        }
        for (int c = 0; c < 3; c++) { W.k.frc[i][c] = acc.w[c]; W.k.frc[i][3 + c] = acc.v[c]; }
    }
    __syncwarp();
    for (int i = 0; i < nb; i++) {
        const int p = m.b_parent[i];
        if (lane < 6) W.k.frc[i][lane] += p >= 0 ? W.k.frc[p][lane] : (lane >= 3 ? -m.g[lane - 3] : 0.0);
        __syncwarp();
    }
    if (lane < nb) {
        const int i = lane;
        Sv6 vi, acc, Ia, Iv, vIv;
        SInert I;
        I.m = W.k.inert[i][0];
        for (int k = 0; k < 3; k++) { I.h[k] = W.k.inert[i][1 + k]; vi.w[k] = W.k.vel[i][k]; vi.v[k] = W.k.vel[i][3 + k]; acc.w[k] = W.k.frc[i][k]; acc.v[k] = W.k.frc[i][3 + k]; }
        for (int k = 0; k < 9; k++) I.I[k] = W.k.inert[i][4 + k];
        inert_apply(Ia, I, acc);
        inert_apply(Iv, I, vi);
        sv_cross_force(vIv, vi, Iv);
        for (int c = 0; c < 3; c++) { W.k.frc[i][c] = Ia.w[c] + vIv.w[c]; W.k.frc[i][3 + c] = Ia.v[c] + vIv.v[c]; }
    }
    __syncwarp();
    for (int i = nb - 1; i >= 0; i--) {
        const int p = m.b_parent[i];
        if (p >= 0 && lane < 6) W.k.frc[p][lane] += W.k.frc[i][lane];
        __syncwarp();
    }
    if (lane < nd) {
        const int b = m.d_body[lane];
        double s = 0;
        for (int c = 0; c < 6; c++) s += W.k.S[lane][c] * W.k.frc[b][c];
        W.bias[lane] = s;
    }
    __syncwarp();
    if (smode == 0) return false;
    STAGE_SYNC(2);   // 2: RNE done
    // ---- composite inertias (lanes = 13 components), joint-space inertia (lane = dof)
    for (int i = nb - 1; i >= 0; i--) {
        const int p = m.b_parent[i];
        if (p >= 0 && lane < 13) W.k.inert[p][lane] += W.k.inert[i][lane];
        __syncwarp();
    }
    for (int e = lane; e < NTRI; e += 32) W.M[e] = 0;
    __syncwarp();
    if (lane < nd) {
        const int i = lane, b = m.d_body[i];
        SInert I;
        Sv6 s, F;
        I.m = W.k.inert[b][0];
        for (int k = 0; k < 3; k++) { I.h[k] = W.k.inert[b][1 + k]; s.w[k] = W.k.S[i][k]; s.v[k] = W.k.S[i][3 + k]; }
        for (int k = 0; k < 9; k++) I.I[k] = W.k.inert[b][4 + k];
        inert_apply(F, I, s);
        for (int j = i; j >= 0; j = m.d_parent[j]) {
            double dsum = 0;
            for (int c = 0; c < 3; c++) dsum += W.k.S[j][c] * F.w[c];
            for (int c = 0; c < 3; c++) dsum += W.k.S[j][3 + c] * F.v[c];
            if (j == i) dsum += m.d_armature[i];
            W.M[TRI(i, j)] = dsum;   // j <= i along the ancestor chain: lower triangle
        }
    }
    __syncwarp();
    // ---- forces
    if (lane < nd) W.tau[lane] = -m.d_damping[lane] * W.qd[lane] - W.bias[lane] + (((comp >> lane) & 1u) ? W.bias_prev[lane] : 0.0);
    __syncwarp();
    if (lane < m.nact) {
        const int a = lane, k = m.a_dof[a];
        double c = W.ctrl[a];
        if (m.a_ctrllimited[a]) c = c < m.a_ctrlrange[a][0] ? m.a_ctrlrange[a][0] : (c > m.a_ctrlrange[a][1] ? m.a_ctrlrange[a][1] : c);
        const double qq = m.d_qadr[k] >= 0 ? W.q[m.d_qadr[k]] : 0.0;
        double f;
        if (m.a_kind[a] == 1) f = m.a_kp[a] * c - m.a_kp[a] * (m.a_gear[a] * qq);
        else if (m.a_kind[a] == 2) f = m.a_kv[a] * c - m.a_kv[a] * (m.a_gear[a] * W.qd[k]);
        else f = c;
        if (m.a_forcelimited[a]) f = f < m.a_forcerange[a][0] ? m.a_forcerange[a][0] : (f > m.a_forcerange[a][1] ? m.a_forcerange[a][1] : f);
        W.tau[k] += m.a_gear[a] * f;   // one actuator per dof in the scenes compiled here
    }
    __syncwarp();
    STAGE_SYNC(3);   // 3: inertia matrix and forces done
    if (lane < nd) W.qacc0[lane] = W.tau[lane];
    __syncwarp();
    // upper half-warp: the factor the velocity update needs (Euler: M + h * diag(damping); explicit RK4 stage: M itself)
    if (W.blk0 == SAWYER_BLK0) w_factor_solve<SAWYER_BLK0>(W.M, nullptr, 0.0, nd, W.L, W.invd, W.qacc0, W.col2, lane, W.L2, W.invd2, m.d_damping, smode == 2 ? 0.0 : m.h);
    else if (nd <= SMALL_ND) w_factor_solve<SMALL_ND>(W.M, nullptr, 0.0, nd, W.L, W.invd, W.qacc0, W.col2, lane, W.L2, W.invd2, m.d_damping, smode == 2 ? 0.0 : m.h);
    else w_factor_solve<0>(W.M, nullptr, 0.0, nd, W.L, W.invd, W.qacc0, W.col2, lane, W.L2, W.invd2, m.d_damping, smode == 2 ? 0.0 : m.h);

    STAGE_SYNC(4);   // 4: unconstrained acceleration done
    return true;
}

template <int WB, int WG, int WC, bool RK, int ND>
__device__ __noinline__ int w_stage_contacts(const DynDev &m, const DynDev *__restrict__ mg, WarpWS<WB, WG, WC> &W, int lane, int &nlim_out) {
    const int nd = ND > 0 ? ND : m.nd;
    long long t_last = c_tune.prof ? clock64() : 0;
    // ---- constraint rows.  Limits first (dof order), then contacts (pair order).
    int nlim = 0;
    {
        // each lane inspects one (dof, side); compaction keeps dof-major order
        const int k = lane >> 1, side = lane & 1;
        bool act = false;
        double dist = 0;
        if (k < nd && m.d_limited[k] && m.d_qadr[k] >= 0) {
            const double qq = W.q[m.d_qadr[k]];
            dist = side == 0 ? qq - m.d_range[k][0] : m.d_range[k][1] - qq;
            act = dist < m.d_margin[k];
        }
        const unsigned mask = __ballot_sync(FULL, act);
        nlim = __popc(mask);
        const int r = __popc(mask & ((1u << lane) - 1));
        if (act) {
            for (int j = 0; j < nd; j++) W.Y[r * YS + j] = 0;
            W.Y[r * YS + k] = side == 0 ? 1.0 : -1.0;
        }
        // row parameters are kept in registers of the lane that owns the row: gather them below
        // by recomputing from (k, side) of the r-th set bit
        __syncwarp();
    }
    int ncp = 0;
    if (m.enable_contacts && m.npair > 0) {
        // Lane-indexed model reads go through the global copy `mg` (coalesced, L1 / L2 resident) - the
        // constant bank serialises divergent addresses.  World frames of the contact geoms: centres for all
        // of them, rotation matrices only where they are not already available (static geoms: precomputed
        // in mg->g_mat; moving geoms with an identity local rotation: the body's xmat).
        for (int g = lane; g < m.ngeom; g += 32) {
            const int body = mg->g_body[g];
            if (body < 0) { for (int k = 0; k < 3; k++) W.gpos[g][k] = mg->g_pos[g][k]; }
            else {
                const double *X = W.k.xmat[body];
                const double p0 = mg->g_pos[g][0], p1 = mg->g_pos[g][1], p2 = mg->g_pos[g][2];
                for (int k = 0; k < 3; k++) W.gpos[g][k] = W.k.xpos[body][k] + X[3 * k] * p0 + X[3 * k + 1] * p1 + X[3 * k + 2] * p2;
            }
        }
        if (lane < m.ngm) {
            const int g = m.gm_geom[lane];
            const double *X = W.k.xmat[mg->g_body[g]], *Rl = mg->g_mat[g];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) W.gmat[lane][3 * r + c] = X[3 * r] * Rl[c] + X[3 * r + 1] * Rl[3 + c] + X[3 * r + 2] * Rl[6 + c];
        }
        __syncwarp();
        auto geom_R = [&](int g) -> const double * {
            const int slot = mg->g_mslot[g];
            return slot == -2 ? mg->g_mat[g] : (slot == -1 ? W.k.xmat[mg->g_body[g]] : W.gmat[slot]);
        };
        int ncand = 0;
        for (int base = 0; base < m.npair; base += 32) {
            const int p = base + lane;
            bool keep = false;
            if (p < m.npair) {
                const int a = mg->p_g1[p], b = mg->p_g2[p];
                keep = true;
                const double ma = mg->g_margin[a], mb = mg->g_margin[b], margin = ma > mb ? ma : mb;
                const int ta = mg->g_type[a], tb = mg->g_type[b];
                const double ra = mg->g_rbound[a], rb = mg->g_rbound[b];
                if (ta != 0 && tb != 0) {
                    const double d[3] = {W.gpos[b][0] - W.gpos[a][0], W.gpos[b][1] - W.gpos[a][1], W.gpos[b][2] - W.gpos[a][2]};
                    const double bound = ra + rb + margin;
                    keep = !(d_dot(d, d) > bound * bound);
                    // box pairs: the bounding sphere of a flat box (table, bin walls) is hopelessly loose; test the
                    // other geom's bounding sphere against the box itself (exact point-to-box distance)
                    if (keep && (ta == 6 || tb == 6)) {
                        const bool box_a = ta == 6 && (tb != 6 || ra >= rb);
                        const int gx = box_a ? a : b;
                        const double *R = geom_R(gx), sg = box_a ? 1.0 : -1.0;   // other centre - box centre = sg * d
                        double e2 = 0;
                        for (int k = 0; k < 3; k++) {
                            const double l = sg * (R[k] * d[0] + R[3 + k] * d[1] + R[6 + k] * d[2]), ex = fabs(l) - mg->g_size[gx][k];
                            if (ex > 0) e2 += ex * ex;
                        }
                        const double bo = (box_a ? rb : ra) + margin + 1e-9;
                        keep = !(e2 > bo * bo);
                        // capsule / cylinder against the box: per box axis the gap between the slab and the projected
                        // segment; the root sum of squares bounds the segment-box distance from below
                        const int go = box_a ? b : a, to = box_a ? tb : ta;
                        if (keep && (to == 3 || to == 5)) {
                            const double *Ro = geom_R(go);
                            const double ax[3] = {Ro[2], Ro[5], Ro[8]}, hl = mg->g_size[go][1];
                            double g2 = 0;
                            for (int k = 0; k < 3; k++) {
                                const double l = sg * (R[k] * d[0] + R[3 + k] * d[1] + R[6 + k] * d[2]);
                                const double ak = R[k] * ax[0] + R[3 + k] * ax[1] + R[6 + k] * ax[2];
                                const double gap = fabs(l) - fabs(ak) * hl - mg->g_size[gx][k];
                                if (gap > 0) g2 += gap * gap;
                            }
                            const double bc = mg->g_size[go][0] + margin + 1e-9;
                            keep = !(g2 > bc * bc);
                        }
                    }
                    // capsule / cylinder pairs: both shapes lie inside the capsule (segment, radius) around
                    // their axis, so the segment-segment distance bounds the true distance from below
                    if (keep && (ta == 3 || ta == 5) && (tb == 3 || tb == 5)) {
                        const double *Ra = geom_R(a), *Rb = geom_R(b);
                        const double a1[3] = {Ra[2], Ra[5], Ra[8]}, a2[3] = {Rb[2], Rb[5], Rb[8]};
                        const double h1 = mg->g_size[a][1], h2 = mg->g_size[b][1];
                        const double r[3] = {-d[0], -d[1], -d[2]};
                        const double bb = d_dot(a1, a2), c = d_dot(a1, r), f = d_dot(a2, r), den = 1.0 - bb * bb;
                        double sp = den > 1e-9 ? c_clamp((bb * f - c) / den, -h1, h1) : 0.0, tp = bb * sp + f;
                        if (tp < -h2) { tp = -h2; sp = c_clamp(bb * tp - c, -h1, h1); }
                        else if (tp > h2) { tp = h2; sp = c_clamp(bb * tp - c, -h1, h1); }
                        double w[3];
                        for (int k = 0; k < 3; k++) w[k] = r[k] + sp * a1[k] - tp * a2[k];
                        const double lo = sqrt(d_dot(w, w)) - mg->g_size[a][0] - mg->g_size[b][0];
                        keep = lo < margin + 1e-9;
                    }
                } else {
                    // plane pairs: nothing of geom b is closer to the plane than its centre height minus rbound
                    const int gp = ta == 0 ? a : b, go = ta == 0 ? b : a;
                    const double *Rp = geom_R(gp);
                    const double hgt = (W.gpos[go][0] - W.gpos[gp][0]) * Rp[2] + (W.gpos[go][1] - W.gpos[gp][1]) * Rp[5] +
                                       (W.gpos[go][2] - W.gpos[gp][2]) * Rp[8];
                    keep = hgt - (ta == 0 ? rb : ra) < margin + 1e-9;
                }
            }
            const unsigned mask = __ballot_sync(FULL, keep);
            const int pos = ncand + __popc(mask & ((1u << lane) - 1));
            if (keep && pos < WCAND) W.cand[pos] = p;
            ncand += __popc(mask);
        }
        if (ncand > WCAND) ncand = WCAND;
        __syncwarp();
        PROF_MARK(23);
        if (c_tune.prof && lane == 0) atomicAdd(&g_prof[29], (unsigned long long)ncand);
        // ---- narrow phase, lane = surviving pair.  Geometry is referenced in place (shared / global memory):
        // nothing here may live in a per-thread stack array, a single lane walking local memory drags whole
        // 32-lane-interleaved lines through L1.  Pairs whose contact needs the polygon clipper (box face
        // contacts) run that part one lane at a time on the warp's shared scratch; contact points are
        // committed in pair order.
        const int maxcp = min(WC / 3, (WC - nlim) / 3);
        for (int base = 0; base < ncand && ncp < maxcp; base += 32) {
            const int ci = base + lane;
            CPoint cps[4];
            int nc = 0, a = 0, b = 0, clip_code = -1;
            double margin = 0;
            if (ci < ncand) {
                const int p = W.cand[ci];
                a = mg->p_g1[p]; b = mg->p_g2[p];
                const double ma = mg->g_margin[a], mb = mg->g_margin[b];
                margin = ma > mb ? ma : mb;
                const int ta = mg->g_type[a], tb = mg->g_type[b];
                CGeom ga{W.gpos[a], geom_R(a), mg->g_size[a], ta}, gb{W.gpos[b], geom_R(b), mg->g_size[b], tb};
                if (ta == 6 && tb == 6) {
                    int code = 0;
                    double depth = 0;
                    const int kind = box_box_sat(ga, gb, margin, code, depth);
                    if (kind == 1) nc = box_box_edge(cps, ga, gb, code, depth);
                    else if (kind == 2) clip_code = code;
                } else
                    nc = pair_contacts_ni(cps, ga, gb, margin);
            }
            unsigned pending = __ballot_sync(FULL, nc > 0 || clip_code >= 0);
            while (pending && ncp < maxcp) {
                const int l = __ffs(pending) - 1;
                pending &= pending - 1;
                int ncl = 0;
                // clipper output + scratch: the last rows of W.Y (constraint rows are written after the narrow phase;
                // the limit rows already there occupy at most the first nd <= 16 rows)
                double *scr = W.Y + (WC - 6) * YS;
                static_assert(6 * YS >= 28 + 24 && WC - 6 >= WD, "clipper scratch does not fit behind the limit rows");
                CPoint *cout = reinterpret_cast<CPoint *>(scr);
                const int code_l = __shfl_sync(FULL, clip_code, l);
                int nclip = 0;
                if (code_l >= 0) {   // box face contact of lane l's pair: all lanes clip the incident face together
                    const int a_l = __shfl_sync(FULL, a, l), b_l = __shfl_sync(FULL, b, l);
                    const double margin_l = shfl_d(margin, l);
                    CGeom ga{W.gpos[a_l], geom_R(a_l), mg->g_size[a_l], 6}, gb{W.gpos[b_l], geom_R(b_l), mg->g_size[b_l], 6};
                    nclip = box_box_face_warp(cout, ga, gb, code_l, margin_l, reinterpret_cast<double(*)[3]>(scr + 28), lane);
                }
                if (lane == l) {
                    const CPoint *src = cps;
                    if (clip_code >= 0) { nc = nclip; src = cout; }
                    const double fa = mg->g_friction[a][0], fb = mg->g_friction[b][0];
                    for (int qn = 0; qn < nc; qn++) {
                        const int slot = ncp + qn;
                        if (slot >= maxcp) break;
                        for (int k = 0; k < 3; k++) { W.cpos[slot][k] = src[qn].pos[k]; W.cn[slot][k] = src[qn].n[k]; }
                        W.cdist[slot] = src[qn].dist;
                        W.cmargin[slot] = margin;
                        W.cmu[slot] = fa > fb ? fa : fb;
                        W.cga[slot] = a; W.cgb[slot] = b;
                        W.csig[slot] = W.cand[ci] * 16 + qn * 4;
                    }
                    ncl = nc;
                }
                ncp += __shfl_sync(FULL, ncl, l);
                if (ncp > maxcp) ncp = maxcp;
                __syncwarp();
            }
        }
        __syncwarp();
        PROF_MARK(24);
    }
    nlim_out = nlim;
    return ncp;
}

template <int WB, int WG, int WC, bool RK, int ND>
__device__ __noinline__ void w_stage_solve(const DynDev &m, const DynDev *__restrict__ mg, WarpWS<WB, WG, WC> &W, int smode, int lane, int nlim, int ncp, int &nwt_out, double &cforce_out, bool sync) {
    const int nb = m.nb, nd = ND > 0 ? ND : m.nd;
    long long t_last = c_tune.prof ? clock64() : 0;
    STAGE_SYNC(5);   // 5: contact points done
    const int nc = nlim + 3 * ncp;
    double fcv = 0.0;  // lane k: constraint force on dof k
    if (nc > 0) {
        // ---- lane = row: parameters, Jacobian (contacts), reference acceleration
        const int r = lane;
        int type = -1, sig = 0;
        double pos = 0, margin = 0, mu = 0, solref[2] = {0.02, 1}, solimp[5] = {0.9, 0.95, 0.001, 0.5, 2};
        if (r < nlim) {
            // recover (dof, side) of the r-th active limit in dof-major order
            int cnt = 0;
            for (int k = 0; k < nd; k++) {
                if (!m.d_limited[k] || m.d_qadr[k] < 0) continue;
                const double qq = W.q[m.d_qadr[k]];
                for (int side = 0; side < 2; side++) {
                    const double dist = side == 0 ? qq - m.d_range[k][0] : m.d_range[k][1] - qq;
                    if (dist < m.d_margin[k]) {
                        if (cnt == r) {
                            type = 0; pos = dist; margin = m.d_margin[k]; sig = -(2 * k + side + 1);
                            solref[0] = m.d_solref[k][0]; solref[1] = m.d_solref[k][1];
                            for (int j = 0; j < 5; j++) solimp[j] = m.d_solimp[k][j];
                        }
                        cnt++;
                    }
                }
            }
        } else if (r < nc) {
            const int c = (r - nlim) / 3, dirn = (r - nlim) % 3;
            type = dirn == 0 ? 1 : 2;
            sig = W.csig[c] + dirn;
            pos = W.cdist[c]; margin = W.cmargin[c]; mu = W.cmu[c];
            {   // pair parameters: mean of the two geoms (solmix 1 : 1)
                const int ga_ = W.cga[c], gb_ = W.cgb[c];
                for (int j = 0; j < 2; j++) solref[j] = 0.5 * (mg->g_solref[ga_][j] + mg->g_solref[gb_][j]);
                for (int j = 0; j < 5; j++) solimp[j] = 0.5 * (mg->g_solimp[ga_][j] + mg->g_solimp[gb_][j]);
            }
            double n[3] = {W.cn[c][0], W.cn[c][1], W.cn[c][2]}, t1[3], t2[3], ref[3] = {0, 0, 0}, cp[3] = {W.cpos[c][0], W.cpos[c][1], W.cpos[c][2]};
            ref[fabs(n[0]) < 0.7 ? 0 : 1] = 1.0;
            d_cross(t1, n, ref);
            const double l = sqrt(d_dot(t1, t1));
            for (int k = 0; k < 3; k++) t1[k] /= l;
            d_cross(t2, n, t1);
            const double *dir = dirn == 0 ? n : (dirn == 1 ? t1 : t2);
            for (int j = 0; j < nd; j++) W.Y[r * YS + j] = 0;
            for (int side = 0; side < 2; side++) {
                int body = m.g_body[side ? W.cgb[c] : W.cga[c]];
                const double sg = side ? 1.0 : -1.0;
                while (body >= 0 && m.b_jtype[body] < 0) body = m.b_parent[body];
                if (body < 0) continue;
                for (int k = m.b_dadr[body] + (m.b_jtype[body] == 0 ? 5 : 0); k >= 0; k = m.d_parent[k]) {
                    double t[3], j3[3], sw[3] = {W.k.S[k][0], W.k.S[k][1], W.k.S[k][2]};
                    d_cross(t, sw, cp);
                    for (int cc = 0; cc < 3; cc++) j3[cc] = W.k.S[k][3 + cc] + t[cc];
                    W.Y[r * YS + k] += sg * d_dot(dir, j3);
                }
            }
        }
        double jv = 0;
        if (r < nc) {
            double jw = 0;
            for (int k = 0; k + 1 < nd; k += 2) { jv += W.Y[r * YS + k] * W.qd[k]; jw += W.Y[r * YS + k + 1] * W.qd[k + 1]; }
            if (nd & 1) jv += W.Y[r * YS + nd - 1] * W.qd[nd - 1];
            jv += jw;
        }
        __syncwarp();   // every Jacobian row is complete: the kinematic arrays may now be overwritten (W.A aliases them)
        // The Hessian M + J^T K J stays block diagonal (arm | free object) unless a constraint row touches dofs of both trees
        // (the arm in contact with the object): only then the dense elimination is needed.
        bool both = false;
        if (W.blk0 > 0 && r < nc) {
            bool lo_nz = false, hi_nz = false;
            for (int k = 0; k < nd; k++) { const bool nz = W.Y[r * YS + k] != 0.0; if (k < W.blk0) lo_nz |= nz; else hi_nz |= nz; }
            both = lo_nz && hi_nz;
        }
        const int hblk = __any_sync(FULL, both) ? 0 : W.blk0;
        // ---- regulariser R = (1 - imp) / imp * (J M^-1 J^T)_rr with y = L^-1 J^T by forward substitution (lane = row)
        double aref = 0, Dr = 0;
        if (r < nc) {
            const double *jr = W.Y + r * YS;
            double *yr = W.A + r * YS;
            int k0 = 0;
            while (k0 < nd && jr[k0] == 0.0) k0++;          // leading exact zeros stay zero
            double diag = 0;
            for (int k = k0; k < nd; k++) {
                double s = jr[k];
                const int j0 = W.blk[k] > k0 ? W.blk[k] : k0;   // L[k][j] = 0 outside k's tree
                double s1 = 0;   // two accumulators: halves the dependent-add chain of the substitution
                int j = j0;
                for (; j + 1 < k; j += 2) { s -= W.L[TRI(k, j)] * yr[j]; s1 -= W.L[TRI(k, j + 1)] * yr[j + 1]; }
                if (j < k) s -= W.L[TRI(k, j)] * yr[j];
                s = (s + s1) * W.invd[k];
                yr[k] = s;
                diag += s * s;
            }
            double K, B, imp;
            kbi_ni(m, solref, solimp, pos, margin, K, B, imp);
            aref = type <= 1 ? (-B * jv - K * imp * (pos - margin)) : (-B * jv);
            double Rg = (1 - imp) / imp * diag;
            if (Rg < DYN_MINVAL) Rg = DYN_MINVAL;
            Dr = 1.0 / Rg;
        }
        // friction rows share the normal row's D (impratio 1); `base` = first row of this lane's block
        const int dirn = type == 2 ? ((sig & 3)) : 0, base = r - dirn;
        Dr = shfl_d(Dr, base & 31);
        PROF_MARK(25);
        // ---- Newton solver on MuJoCo's primal problem (mj_solNewton; same algorithm as oracle/orc_dyn.c):
        //   min_a 1/2 (a - a0)^T M (a - a0) + sum_c s_c(J_c a - aref_c),  elliptic cones, exact line search.
        // lane = row for x = J a - aref, the cone terms and J p; lane = dof for M a, the gradient and the search direction.
        const int ntri = nd * (nd + 1) / 2;
        double trM = (lane < nd) ? W.M[TRI(lane, lane)] : 0.0;
        trM = warp_sum(trM);
        const double scale = 1.0 / (trM > DYN_MINVAL ? trM : DYN_MINVAL);
        double cost = 0, x0 = 0, x1 = 0, x2 = 0, gsr = 0, h0 = 0, h1 = 0, h2 = 0;
        const double Dm = Dr / (1 + mu * mu);   // middle-zone weight of this lane's contact
        // evaluates cost / gradient (W.rhs) / M a - tau (W.z) at W.a
        auto newton_eval = [&]() {
            double mat = 0, cq = 0;
            if (lane < nd) {
                double m1 = 0;   // two accumulators: the dependent-add chain is the latency of these loops
                for (int j = 0; j + 1 < nd; j += 2) {
                    mat += W.M[j <= lane ? TRI(lane, j) : TRI(j, lane)] * W.a[j];
                    m1 += W.M[j + 1 <= lane ? TRI(lane, j + 1) : TRI(j + 1, lane)] * W.a[j + 1];
                }
                if (nd & 1) mat += W.M[TRI(nd - 1, lane)] * W.a[nd - 1];
                mat += m1;
                mat -= W.tau[lane];
                cq = 0.5 * (W.a[lane] - W.qacc0[lane]) * mat;
                W.z[lane] = mat;
            }
            double x = 0;
            if (r < nc) {
                double xb = 0;
                for (int k = 0; k + 1 < nd; k += 2) { x += W.Y[r * YS + k] * W.a[k]; xb += W.Y[r * YS + k + 1] * W.a[k + 1]; }
                if (nd & 1) x += W.Y[r * YS + nd - 1] * W.a[nd - 1];
                x = (x + xb) - aref;
            }
            x0 = shfl_d(x, base & 31); x1 = shfl_d(x, (base + 1) & 31); x2 = shfl_d(x, (base + 2) & 31);
            double sc = 0;
            gsr = 0; h0 = 0; h1 = 0; h2 = 0;
            if (type == 0) {
                if (x < 0) { gsr = Dr * x; h0 = Dr; sc = 0.5 * Dr * x * x; }
            } else if (type >= 1) {
                const double t = sqrt(x1 * x1 + x2 * x2);
                if (x0 >= mu * t) { }                                   // top zone: inactive
                else if (mu * x0 + t <= 0) {                            // bottom zone: quadratic in all three rows
                    gsr = Dr * x; 
                    if (dirn == 0) { h0 = Dr; sc = 0.5 * Dr * (x0 * x0 + x1 * x1 + x2 * x2); } else if (dirn == 1) h1 = Dr; else h2 = Dr;
                } else {                                                // middle zone: distance to the cone surface
                    // one reciprocal instead of six fp64 divisions (n1, n2 = unit tangential direction)
                    const double it = 1.0 / t, n1 = x1 * it, n2 = x2 * it, e = x0 - mu * t, u1 = -mu * n1, u2 = -mu * n2;
                    const double c = -Dm * e * mu * it;
                    if (dirn == 0) { gsr = Dm * e; h0 = Dm; h1 = Dm * u1; h2 = Dm * u2; sc = 0.5 * Dm * e * e; }
                    else if (dirn == 1) { gsr = Dm * e * u1; h0 = Dm * u1; h1 = Dm * u1 * u1 + c * (1 - n1 * n1); h2 = Dm * u1 * u2 + c * (-n1 * n2); }
                    else { gsr = Dm * e * u2; h0 = Dm * u2; h1 = Dm * u1 * u2 + c * (-n1 * n2); h2 = Dm * u2 * u2 + c * (1 - n2 * n2); }
                }
            }
            cost = warp_sum(cq + sc);
            if (lane < WC) W.f[lane] = gsr;
            __syncwarp();
            if (lane < nd) {
                double gg = mat, g1 = 0;
                for (int s2 = 0; s2 + 1 < nc; s2 += 2) { gg += W.Y[s2 * YS + lane] * W.f[s2]; g1 += W.Y[(s2 + 1) * YS + lane] * W.f[s2 + 1]; }
                if (nc & 1) gg += W.Y[(nc - 1) * YS + lane] * W.f[nc - 1];
                W.rhs[lane] = gg + g1;
            }
            __syncwarp();
        };
        // Hessian (lower triangle) -> W.L from the cone terms of the latest evaluation
        auto newton_hess = [&]() {
            // K_r = sum_q Hc[dirn][q] J_(base+q)  ->  H = M + sum_r J_r^T K_r
            if (r < nc) {
                const bool blockrow = type >= 1;
                for (int j = 0; j < nd; j++) {
                    double kk = h0 * W.Y[base * YS + j];
                    if (blockrow) kk += h1 * W.Y[(base + 1) * YS + j] + h2 * W.Y[(base + 2) * YS + j];
                    W.A[r * YS + j] = kk;
                }
            }
            __syncwarp();
            for (int e = lane; e < ntri; e += 32) {
                int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
                while (TRI(i + 1, 0) <= e) i++;
                while (TRI(i, 0) > e) i--;
                const int j = e - TRI(i, 0);
                double hh = W.M[e], hb = 0;
                for (int s2 = 0; s2 + 1 < nc; s2 += 2) { hh += W.Y[s2 * YS + i] * W.A[s2 * YS + j]; hb += W.Y[(s2 + 1) * YS + i] * W.A[(s2 + 1) * YS + j]; }
                if (nc & 1) hh += W.Y[(nc - 1) * YS + i] * W.A[(nc - 1) * YS + j];
                W.L[e] = hh + hb;
            }
            __syncwarp();
        };
        // warm start (mj_solNewton / mj_warmstart): the previous substep's acceleration unless the unconstrained one costs
        // less.  Both candidates converge to the same minimiser, but only to within the solver tolerance: starting where
        // the reference starts keeps kernel and oracle on the same iteration path over long horizons.
        // The Hessian is only assembled once it is known that a Newton step follows.
        bool have_eval = false;
        if (W.wn) {
            // cost at the unconstrained acceleration, cheaply: M qacc0 = tau, so only the constraint penalties remain
            double xq = 0;
            if (r < nc) {
                double xb = 0;
                for (int k = 0; k + 1 < nd; k += 2) { xq += W.Y[r * YS + k] * W.qacc0[k]; xb += W.Y[r * YS + k + 1] * W.qacc0[k + 1]; }
                if (nd & 1) xq += W.Y[r * YS + nd - 1] * W.qacc0[nd - 1];
                xq = (xq + xb) - aref;
            }
            const double q0 = shfl_d(xq, base & 31), q1 = shfl_d(xq, (base + 1) & 31), q2 = shfl_d(xq, (base + 2) & 31);
            double sc0 = 0;
            if (type == 0) { if (xq < 0) sc0 = 0.5 * Dr * xq * xq; }
            else if (type == 1) {
                const double t = sqrt(q1 * q1 + q2 * q2);
                if (q0 >= mu * t) { }
                else if (mu * q0 + t <= 0) sc0 = 0.5 * Dr * (q0 * q0 + q1 * q1 + q2 * q2);
                else { const double e = q0 - mu * t; sc0 = 0.5 * Dm * e * e; }
            }
            const double c0 = warp_sum(sc0);
            if (lane < nd) W.a[lane] = W.wa[lane];
            __syncwarp();
            newton_eval();
            if (cost < c0) have_eval = true;
            else { if (lane < nd) W.a[lane] = W.qacc0[lane]; __syncwarp(); }
        } else {
            if (lane < nd) W.a[lane] = W.qacc0[lane];
            __syncwarp();
        }
        int it = 0;
        bool first = true;
        double old = 0;
        for (;;) {
            if (!have_eval) newton_eval();
            have_eval = false;
            if (!first && scale * (old - cost) < m.tolerance) break;
            first = false;
            if (it++ >= m.iterations) break;
            const double gme = lane < nd ? W.rhs[lane] : 0.0;
            const double gn = warp_sum(gme * gme);
            if (scale * sqrt(gn) < m.tolerance) break;
            newton_hess();
            if (c_tune.prof && lane == 0) { atomicAdd(&g_prof[20], 1ULL); if (it == m.iterations) atomicAdd(&g_prof[28], 1ULL); }
            nwt_out += 1;
            // search direction p = -H^-1 g (W.rhs in place; W.z keeps M a - tau)
            const double matme = lane < nd ? W.z[lane] : 0.0;
            if (lane < nd) W.rhs[lane] = -gme;
            __syncwarp();
            if (hblk == SAWYER_BLK0) w_factor_solve<SAWYER_BLK0>(W.L, nullptr, 0.0, nd, W.L, W.invd, W.rhs, W.col2, lane);
            else if (nd <= SMALL_ND) w_factor_solve<SMALL_ND>(W.L, nullptr, 0.0, nd, W.L, W.invd, W.rhs, W.col2, lane);
            else w_factor_solve<0>(W.L, nullptr, 0.0, nd, W.L, W.invd, W.rhs, W.col2, lane);
            double pme = 0, mp = 0;
            if (lane < nd) {
                pme = W.rhs[lane];
                double mq = 0;
                for (int j = 0; j + 1 < nd; j += 2) {
                    mp += W.M[j <= lane ? TRI(lane, j) : TRI(j, lane)] * W.rhs[j];
                    mq += W.M[j + 1 <= lane ? TRI(lane, j + 1) : TRI(j + 1, lane)] * W.rhs[j + 1];
                }
                if (nd & 1) mp += W.M[TRI(nd - 1, lane)] * W.rhs[nd - 1];
                mp += mq;
            }
            const double q1 = warp_sum(pme * matme), q2 = warp_sum(pme * mp), d0 = warp_sum(pme * gme);
            double jp = 0;
            if (r < nc)
            {
                double jq = 0;
                for (int k = 0; k + 1 < nd; k += 2) { jp += W.Y[r * YS + k] * W.rhs[k]; jq += W.Y[r * YS + k + 1] * W.rhs[k + 1]; }
                if (nd & 1) jp += W.Y[r * YS + nd - 1] * W.rhs[nd - 1];
                jp += jq;
            }
            const double jp0 = shfl_d(jp, base & 31), jp1 = shfl_d(jp, (base + 1) & 31), jp2 = shfl_d(jp, (base + 2) & 31);
            // exact line search: safeguarded Newton iteration on phi'(alpha)
            double alpha = 1.0, lo = 0.0, hi = -1.0;
            for (int ls = 0; ls < 24; ls++) {
                double d1 = 0, d2 = 0;
                if (type == 0) {
                    const double xa = x0 + alpha * jp0;
                    if (xa < 0) { d1 = Dr * xa * jp0; d2 = Dr * jp0 * jp0; }
                } else if (type == 1) {
                    const double a0 = x0 + alpha * jp0, a1 = x1 + alpha * jp1, a2 = x2 + alpha * jp2, t = sqrt(a1 * a1 + a2 * a2);
                    if (a0 >= mu * t) { }
                    else if (mu * a0 + t <= 0) { d1 = Dr * (a0 * jp0 + a1 * jp1 + a2 * jp2); d2 = Dr * (jp0 * jp0 + jp1 * jp1 + jp2 * jp2); }
                    else {
                        const double it = 1.0 / t, e = a0 - mu * t, u1 = -mu * a1 * it, u2 = -mu * a2 * it, c = -Dm * e * mu * it;
                        const double uj = jp0 + u1 * jp1 + u2 * jp2, tj = (a1 * jp1 + a2 * jp2) * it;
                        d1 = Dm * e * uj;
                        d2 = Dm * uj * uj + c * (jp1 * jp1 + jp2 * jp2 - tj * tj);
                    }
                }
                d1 = warp_sum(d1) + q1 + alpha * q2;
                d2 = warp_sum(d2) + q2;
                if (fabs(d1) <= 1e-6 * fabs(d0)) break;
                if (d1 < 0) lo = alpha; else hi = alpha;
                double an = alpha - d1 / d2;
                if (hi >= 0) { if (!(an > lo && an < hi)) an = 0.5 * (lo + hi); }
                else if (!(an > lo)) an = 2 * alpha;
                alpha = an;
            }
            if (lane < nd) W.a[lane] += alpha * pme;
            __syncwarp();
            old = cost;
        }
        if (c_tune.prof && lane == 0) { atomicAdd(&g_prof[21], (unsigned long long)nc); atomicAdd(&g_prof[22], 1ULL); }
        PROF_MARK(26);
        // constraint force on the dofs: J^T f = -(J^T grad s) = (M a - tau) - g
        if (lane < nd) { fcv = W.z[lane] - W.rhs[lane]; W.wa[lane] = W.a[lane]; }
        cforce_out = warp_sum(type >= 1 ? fabs(gsr) : 0.0);   // BaseEnv.get_contact_force: sum of |contact-frame force components|
        if (lane == 0) W.wn = 1;
        __syncwarp();
    }
    else {
        cforce_out = 0.0;
        if (lane == 0) W.wn = 0;
    }
    STAGE_SYNC(6);   // 6: constraint forces done
    if (lane < nd) W.rhs[lane] = W.tau[lane] + fcv;
    __syncwarp();
    if (RK && smode == 2) {   // explicit stage of mj_RungeKutta: qacc = M^-1 (tau + J^T f), joint damping is part of tau
        w_solve_stored(W.L2, W.invd2, nd, W.rhs, lane, W.blk0);
        STAGE_SYNC(7);
        return;
    }
    // ---- semi-implicit Euler with implicit joint damping (factor of M + h * diag(damping) from stage 4)
    w_solve_stored(W.L2, W.invd2, nd, W.rhs, lane, W.blk0);
    if (lane < nd) {
        W.qd[lane] += m.h * W.rhs[lane];
        W.v[m.d_vadr[lane]] = W.qd[lane];
    }
    __syncwarp();
    w_integrate_pos(m, W, m.h, lane);
    STAGE_SYNC(7);   // 7: state advanced
}

template <int WB, int WG, int WC, bool RK = false, int ND = 0, int NB = 0>   // RK: the instantiation that also serves smode 2 / 3 (Pusher); the Sawyer kernels keep their code
__device__ __forceinline__ void w_substep(const DynDev &m, const DynDev *__restrict__ mg, WarpWS<WB, WG, WC> &W, unsigned comp, int smode, int lane, int &ncon_out, int &nwt_out, double &cforce_out, const int4 keep_bodies,
                                          bool active, bool sync) {
    if (!active) {   // barrier-only participant: one barrier per stage boundary
        long long t_last = c_tune.prof ? clock64() : 0;
        for (int k = 1; k <= 7; k++) STAGE_SYNC(k);
        return;
    }
    if (!w_stage_dynamics<WB, WG, WC, ND, NB>(m, mg, W, comp, smode, lane, keep_bodies, sync)) return;
    int nlim = 0;
    const int ncp = w_stage_contacts<WB, WG, WC, RK, ND>(m, mg, W, lane, nlim);
    ncon_out = ncp;
    if (RK && smode == 3) return;
    w_stage_solve<WB, WG, WC, RK, ND>(m, mg, W, smode, lane, nlim, ncp, nwt_out, cforce_out, sync);
}

// kept frame slots: 0 = end-effector body, 1 = cube, 2 = right claw, 3 = left claw
template <int WB, int WG, int WC>
__device__ __forceinline__ void w_site(double *out, const WarpWS<WB, WG, WC> &W, int slot, const double *local) {
    double t[3];
    d_mv(t, W.kxmat[slot], local);
    for (int k = 0; k < 3; k++) out[k] = W.kxpos[slot][k] + t[k];
}

template <int WB, int WG, int WC>
__device__ void w_write_obs(const mopa_sawyer_task &T, const WarpWS<WB, WG, WC> &W, float *obs, int lane) {
    if (lane != 0) return;
    int o = 0;
    for (int k = 0; k < 7; k++) obs[o++] = (float)W.q[T.arm_qadr[k]];
    for (int k = 0; k < 7; k++) obs[o++] = (float)W.v[T.arm_vadr[k]];
    for (int k = 0; k < 2; k++) obs[o++] = (float)W.q[T.grip_qadr[k]];
    for (int k = 0; k < 2; k++) obs[o++] = (float)W.v[T.grip_vadr[k]];
    double eef[3];
    w_site(eef, W, 0, T.site_grip);
    for (int k = 0; k < 3; k++) obs[o++] = (float)eef[k];
    const double *eq = W.kxquat[0];
    obs[o++] = (float)eq[1]; obs[o++] = (float)eq[2]; obs[o++] = (float)eq[3]; obs[o++] = (float)eq[0];
    if (T.kind == 2) {   // SawyerAssemblyObstacleEnv._get_obs (:53-59): hole, pegHead, pegEnd, peg_quat (wxyz)
        double hole[3], head[3], end[3];
        w_site(hole, W, 2, T.site_hole);
        w_site(head, W, 1, T.site_right_eef);
        w_site(end, W, 1, T.site_left_eef);
        for (int k = 0; k < 3; k++) obs[o++] = (float)hole[k];
        for (int k = 0; k < 3; k++) obs[o++] = (float)head[k];
        for (int k = 0; k < 3; k++) obs[o++] = (float)end[k];
        for (int k = 0; k < 4; k++) obs[o++] = (float)W.kxquat[1][k];
        obs[o++] = 0.0f; obs[o++] = 0.0f;   // 38 observation floats, row padded to 40
        return;
    }
    if (T.kind == 1) {   // SawyerLiftObstacleEnv._get_obs (:150-161): cube_pos, cube_quat (xyzw), gripper_to_cube = grip_site - cube
        const double *cube = W.kxpos[1], *cq = W.kxquat[1];
        for (int k = 0; k < 3; k++) obs[o++] = (float)cube[k];
        obs[o++] = (float)cq[1]; obs[o++] = (float)cq[2]; obs[o++] = (float)cq[3]; obs[o++] = (float)cq[0];
        for (int k = 0; k < 3; k++) obs[o++] = (float)(eef[k] - cube[k]);
        while (o < 40) obs[o++] = 0.0f;   // 35 observation floats, row padded to 40
        return;
    }
    const double target[3] = {T.target_base[0] + W.q[T.target_qadr[0]], T.target_base[1] + W.q[T.target_qadr[1]], T.target_base[2]};
    for (int k = 0; k < 3; k++) obs[o++] = (float)target[k];
    const double *cube = W.kxpos[1], *cq = W.kxquat[1];
    for (int k = 0; k < 3; k++) obs[o++] = (float)cube[k];
    obs[o++] = (float)cq[1]; obs[o++] = (float)cq[2]; obs[o++] = (float)cq[3]; obs[o++] = (float)cq[0];
    for (int k = 0; k < 3; k++) obs[o++] = (float)(eef[k] - cube[k]);
    for (int k = 0; k < 2; k++) obs[o++] = (float)(cube[k] - target[k]);
}


template <int WB, int WG, int WC, int ENV_WARPS, int ND, int NB>
__global__ void __launch_bounds__(ENV_WARPS * 32, (ENV_WARPS <= 7 && WB <= 14 ? 2 : 1))
env_step_warp_kernel(int model_slot, const DynDev *__restrict__ mg, mopa_env_buffers B, const float *__restrict__ action,
                     int action_stride, const uint8_t *__restrict__ is_planner, const uint8_t *__restrict__ mask, int n, int forward_only,
                     const int32_t *__restrict__ ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpWS<WB, WG, WC> &W = reinterpret_cast<WarpWS<WB, WG, WC> *>(smem_raw)[warp];
    const DynDev &m = c_models[model_slot];
    const mopa_sawyer_task &T = c_tasks[model_slot];
    const int t = blockIdx.x * ENV_WARPS + warp;
    const int e = t < n ? (ids ? ids[t] : t) : 0;
    const bool live = t < n && (!mask || mask[e]);
    if (!live) {   // still take part in the CTA barriers of the substep loop
        if (!forward_only) {
            int dummy = 0;
            double dummyf = 0;
            for (int s = 0; s < T.nsub; s++) w_substep<WB, WG, WC, false, ND, NB>(m, mg, W, 0u, true, lane, dummy, dummy, dummyf, make_int4(0, 0, 0, 0), false, true);
        }
        return;
    }
    for (int k = lane; k < m.nq; k += 32) W.q[k] = B.qpos[(size_t)e * m.nq + k];
    for (int k = lane; k < m.nv; k += 32) W.v[k] = B.qvel[(size_t)e * m.nv + k];
    if (lane < WD) W.bias_prev[lane] = B.bias_prev[(size_t)e * WD + lane];
    if (lane < DMAXA) W.ctrl[lane] = 0.0;
    if (lane == 0) W.wn = 0;
    if (lane < m.nd) { int r = lane; while (m.d_parent[r] >= 0) r = m.d_parent[r]; W.blk[lane] = r; }
    __syncwarp();
    if (lane == 0) {   // two kinematic trees with contiguous dofs: [0, k) and [k, nd)
        int k = 1;
        while (k < m.nd && W.blk[k] == W.blk[0]) k++;
        bool two = k < m.nd;
        for (int j = k; j < m.nd && two; j++) two = W.blk[j] == W.blk[k];
        W.blk0 = (two && k == SAWYER_BLK0 && m.nd - k <= WD - SAWYER_BLK0 && !((c_tune.sync_mask >> 9) & 1)) ? k : 0;   // (tuning bit 9: force the dense elimination)
    }
    __syncwarp();
    double cforce = 0;
    int ncon = 0, nwt = 0;   // contacts of the last substep; Newton steps of this env.step (cost feedback for the caller's grouping)
    const int4 keep = make_int4(T.body_ee, T.body_cube, T.body_rclaw, T.body_lclaw);
    if (forward_only) {
        w_substep<WB, WG, WC, false, ND, NB>(m, mg, W, 0u, false, lane, ncon, nwt, cforce, keep, true, false);
        if (lane < WD) B.bias_prev[(size_t)e * WD + lane] = lane < m.nd ? W.bias[lane] : 0.0;
        w_write_obs(T, W, B.obs + (size_t)e * 40, lane);
        return;
    }
    const int mode = is_planner ? is_planner[e] : 0;
    const bool planner = mode == 1;
    const bool had_prev = B.has_prev[e] != 0;
    if (lane < 7) {
        const double prev = (!planner || !had_prev) ? W.q[T.arm_qadr[lane]] : B.prev_state[(size_t)e * 7 + lane];
        double a = (double)action[(size_t)e * action_stride + lane];
        if (!planner) a = a * T.ac_scale;
        a = a < -T.ac_scale ? -T.ac_scale : (a > T.ac_scale ? T.ac_scale : a);
        W.ctrl[lane] = prev + a;
    }
    // lift: 8-D action, the last entry moves both finger actuators (SawyerEnv._gripper_format_action: gripper qpos + a)
    if (T.kind == 1 && lane < 2 && mode != 2) W.ctrl[7 + lane] = W.q[T.grip_qadr[lane]] + (double)action[(size_t)e * action_stride + 7];
    __syncwarp();
    unsigned comp = 0;
    for (int k = 0; k < 7; k++) comp |= 1u << T.arm_dof[k];
    if (mode == 2) w_substep<WB, WG, WC, false, ND, NB>(m, mg, W, 0u, false, lane, ncon, nwt, cforce, keep, true, false);
    for (int s = 0; s < T.nsub; s++) {
        w_substep<WB, WG, WC, false, ND, NB>(m, mg, W, comp, true, lane, ncon, nwt, cforce, keep, mode != 2, true);
        if (mode != 2 && lane < m.nd) W.bias_prev[lane] = W.bias[lane];
        __syncwarp();
    }
    // ---- instability guard (BaseEnv._do_simulation, env/base.py:388-400: MuJoCo flags NaN / values beyond mjMAXVAL = 1e10 in
    // qpos / qvel / qacc, mujoco-py raises, the env resets and _after_step terminates the episode with -unstable_penalty).
    // Here: the diverged step is discarded (the state row in HBM is left as it was), the episode terminates, the caller
    // resets the environment (the rollout does so on `done`); observation and frames are those of the untouched state.
    bool unstable = false, corrupt = false;
    if (mode != 2) {
        bool bad = false;
        for (int k = lane; k < m.nq; k += 32) { const double x = W.q[k]; if (!(fabs(x) <= 1e10)) bad = true; }
        for (int k = lane; k < m.nv; k += 32) { const double x = W.v[k]; if (!(fabs(x) <= 1e10)) bad = true; }
        unstable = __any_sync(FULL, bad);
        if (unstable) {
            bool bad2 = false;   // the stored row itself is corrupt (only possible through an external write): nothing to show
            for (int k = lane; k < m.nq; k += 32) { const double x = B.qpos[(size_t)e * m.nq + k]; W.q[k] = x; if (!(fabs(x) <= 1e10)) bad2 = true; }
            for (int k = lane; k < m.nv; k += 32) { const double x = B.qvel[(size_t)e * m.nv + k]; W.v[k] = x; if (!(fabs(x) <= 1e10)) bad2 = true; }
            corrupt = __any_sync(FULL, bad2);
            if (lane == 0) W.wn = 0;
            __syncwarp();
            if (!corrupt) w_substep<WB, WG, WC, false, ND, NB>(m, mg, W, 0u, false, lane, ncon, nwt, cforce, keep, true, false);
            ncon = 0; cforce = 0.0;
        }
    }
    // reward / success (frames of the last substep's start state)
    double reward = 0;
    bool success = false, terminal = false;
    int grasp_bits = 0;
    if (T.kind == 1) {   // SawyerLiftObstacleEnv.compute_reward (:92-148)
        double grip[3], d = 0;
        w_site(grip, W, 0, T.site_grip);
        const double *cube = W.kxpos[1];
        for (int k = 0; k < 3; k++) d += (cube[k] - grip[k]) * (cube[k] - grip[k]);
        const double reach = (1 - tanh(10 * sqrt(d))) * 0.1;
        // has_grasp: the contact list of the last mj_step holds can - left-finger and can - right-finger contacts
        // (a planner-failure step has no fresh contact list: mode 2 runs no mj_step)
        bool tl = false, tr = false;
        if (mode != 2) {
            for (int c = 0; c < ncon; c++) {
                const int ga = W.cga[c], gb = W.cgb[c];
                const int other = ga == T.geom_cube ? gb : (gb == T.geom_cube ? ga : -1);
                if (other < 0) continue;
                for (int k = 0; k < 3; k++) { if (other == T.geom_lfinger[k]) tl = true; if (other == T.geom_rfinger[k]) tr = true; }
            }
            grasp_bits = (tl ? 1 : 0) | (tr ? 2 : 0);
        } else if (B.grasp) {
            // planner failure: compute_reward runs without an mj_step and reads the contact list the previous step left in
            // mjData (rl/mopa_rollouts.py:312) - the touch flags of that list are kept per environment
            const int gb = B.grasp[e];
            tl = (gb & 1) != 0; tr = (gb & 2) != 0;
        }
        const bool grasp = tl && tr;
        const double r_grasp = grasp ? 0.35 : 0.0, z_target = T.bin_z + 0.45;
        double r_lift = 0.0;
        if (grasp) {
            const double zd = z_target - cube[2] > 0.0 ? z_target - cube[2] : 0.0;
            r_lift = 0.35 + (1 - tanh(15 * zd)) * (0.5 - 0.35);
        }
        reward = reach > r_grasp ? reach : r_grasp;
        reward = reward > r_lift ? reward : r_lift;
        if (grasp && fabs(cube[2] - z_target) < 0.05) { reward += T.success_reward; success = true; terminal = true; }
    } else if (T.kind == 2) {   // SawyerAssemblyObstacleEnv.compute_reward (:32-51)
        double head[3], hole[3], bottom[3], d1 = 0, d2 = 0;
        w_site(head, W, 1, T.site_right_eef);
        w_site(hole, W, 2, T.site_hole);
        w_site(bottom, W, 2, T.site_hole_bottom);
        for (int k = 0; k < 3; k++) { d1 += (head[k] - hole[k]) * (head[k] - hole[k]); d2 += (head[k] - bottom[k]) * (head[k] - bottom[k]); }
        d1 = sqrt(d1); d2 = sqrt(d2);
        if (d1 < 0.3) reward += 0.4 * (1 - tanh(15 * d1));
        if (d2 < 0.025) { reward += T.success_reward; success = true; terminal = true; }
    } else {             // SawyerPushObstacleEnv.compute_reward (:71-100)
        double re[3], le[3];
        w_site(re, W, 2, T.site_right_eef);
        w_site(le, W, 3, T.site_left_eef);
        const double *cube = W.kxpos[1];
        const double target[2] = {T.target_base[0] + W.q[T.target_qadr[0]], T.target_base[1] + W.q[T.target_qadr[1]]};
        double dgc = 0;
        for (int k = 0; k < 3; k++) { const double d = cube[k] - 0.5 * (re[k] + le[k]); dgc += d * d; }
        dgc = sqrt(dgc);
        const double dct = sqrt((cube[0] - target[0]) * (cube[0] - target[0]) + (cube[1] - target[1]) * (cube[1] - target[1]));
        if (dct < 0.1) reward += 0.5 * (1 - tanh(5 * dct));
        if (dgc < 0.1) reward += 0.1 * (1 - tanh(10 * dgc));
        if (dct < T.distance_threshold) { reward += T.success_reward; success = true; terminal = true; }
    }
    if (unstable) { reward = -T.unstable_penalty; success = false; terminal = true; grasp_bits = 0; }
    if (corrupt) { for (int k = lane; k < 40; k += 32) B.obs[(size_t)e * 40 + k] = 0.0f; }
    else if (mode != 2) w_write_obs(T, W, B.obs + (size_t)e * 40, lane);
    __syncwarp();
    // _after_step: joint-limit projection (set_state + forward), episode accounting
    bool clipped = false;
    if (lane < m.nd && m.d_limited[lane] && m.d_qadr[lane] >= 0) {
        double &x = W.q[m.d_qadr[lane]];
        if (x < m.d_range[lane][0]) { x = m.d_range[lane][0]; clipped = true; }
        else if (x > m.d_range[lane][1]) { x = m.d_range[lane][1]; clipped = true; }
    }
    clipped = __any_sync(FULL, clipped) && !unstable;
    __syncwarp();
    if (clipped) {
        w_substep<WB, WG, WC, false, ND, NB>(m, mg, W, 0u, false, lane, ncon, nwt, cforce, keep, true, false);
        if (lane < m.nd) W.bias_prev[lane] = W.bias[lane];
        __syncwarp();
    }
    if (!unstable) {
        for (int k = lane; k < m.nq; k += 32) B.qpos[(size_t)e * m.nq + k] = W.q[k];
        for (int k = lane; k < m.nv; k += 32) B.qvel[(size_t)e * m.nv + k] = W.v[k];
        if (lane < WD) B.bias_prev[(size_t)e * WD + lane] = lane < m.nd ? W.bias_prev[lane] : 0.0;
        if (mode != 2 && lane < 7) B.prev_state[(size_t)e * 7 + lane] = W.ctrl[lane];
    }
    if (lane == 0) {
        if (mode != 2) B.has_prev[e] = 1;
        const int len = B.ep_len[e] + 1;
        if (len == T.max_episode_steps) terminal = true;
        B.ep_len[e] = len;
        B.ep_rew[e] += reward;
        B.reward[e] = reward;
        B.done[e] = terminal ? 1 : 0;
        B.success[e] = success ? 1 : 0;
        if (B.ncon) B.ncon[e] = ncon;
        if (B.work && mode != 2) B.work[e] = nwt;
        if (B.cforce) B.cforce[e] = mode != 2 ? cforce : 0.0;
        if (B.grasp && mode != 2) B.grasp[e] = (uint8_t)grasp_bits;
        if (B.unstable) B.unstable[e] = unstable ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// PusherObstacle-v0 (BASELINE configs[0]; env/pusher/pusher_obstacle.py:40-67, 185-284; PID of BaseEnv._get_control,
// env/base.py:200-209): 4 hinge joints under velocity actuators, RK4 with dt = 0.01, int(frame_dt / dt) = 100 mj_steps per
// env.step with the PID law re-evaluated before each of them.  Same warp-per-environment workspace and the same w_substep
// as the Sawyer tasks; its own kernel so that the Sawyer kernels keep their code and register budget.
__device__ __forceinline__ unsigned long long pz_mix(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double pz_uniform(unsigned long long seed, unsigned long long stream, unsigned long long counter, unsigned long long dim) {   // mopa_rl_b200/rng.py uniform01
    unsigned long long x = pz_mix(seed ^ (stream * 0x9E3779B97F4A7C15ULL));
    x = pz_mix(x + ((counter << 8) | dim) * 0xD1342543DE82EF95ULL);
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// one mj_step: Euler, or mj_RungeKutta with N = 4 (A = [[1/2], [0, 1/2], [0, 0, 1]], B = [1/6, 1/3, 1/3, 1/6]; every stage runs the
// full forward dynamics incl. collision and the constraint solver; mjData keeps the frames / contacts of the last stage)
template <int WB, int WG, int WC>
__device__ __forceinline__ void w_mj_step(const DynDev &m, const DynDev *__restrict__ mg, WarpWS<WB, WG, WC> &W, int lane, int &ncon, int &nwt,
                                          double &cforce, const int4 keep, bool active, bool sync) {
    if (m.integrator == 0) { w_substep<WB, WG, WC, true>(m, mg, W, 0u, 1, lane, ncon, nwt, cforce, keep, active, sync); return; }
    if (!active) { for (int i = 0; i < 4; i++) w_substep<WB, WG, WC, true>(m, mg, W, 0u, 2, lane, ncon, nwt, cforce, keep, false, sync); return; }
    const int nd = m.nd, va = lane < nd ? m.d_vadr[lane] : 0;
    const double q0a = lane < m.nq ? W.q[lane] : 0.0, q0b = lane + 32 < m.nq ? W.q[lane + 32] : 0.0;
    const double v0a = lane < m.nv ? W.v[lane] : 0.0, v0b = lane + 32 < m.nv ? W.v[lane + 32] : 0.0;
    double Fv[4], Fa[4];
    auto stage_state = [&](double dv, double da) {   // X = X0 + h (dv, da)
        if (lane < m.nq) W.q[lane] = q0a;
        if (lane + 32 < m.nq) W.q[lane + 32] = q0b;
        if (lane < m.nv) W.v[lane] = v0a;
        if (lane + 32 < m.nv) W.v[lane + 32] = v0b;
        if (lane < nd) W.qd[lane] = dv;
        __syncwarp();
        w_integrate_pos(m, W, m.h, lane);
        if (lane < nd) W.v[va] = W.v[va] + m.h * da;
        __syncwarp();
    };
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i > 0) {
            double dv = 0.0, da = 0.0;
#pragma unroll
            for (int j = 0; j < i; j++) {
                const double a = (j == i - 1) ? (i == 3 ? 1.0 : 0.5) : 0.0;
                dv += a * Fv[j]; da += a * Fa[j];
            }
            stage_state(dv, da);
        }
        Fv[i] = lane < nd ? W.v[va] : 0.0;
        w_substep<WB, WG, WC, true>(m, mg, W, 0u, 2, lane, ncon, nwt, cforce, keep, true, sync);
        Fa[i] = lane < nd ? W.rhs[lane] : 0.0;
        __syncwarp();
    }
    double dv = 0.0, da = 0.0;
    const double Bc[4] = {1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6};
#pragma unroll
    for (int j = 0; j < 4; j++) { dv += Bc[j] * Fv[j]; da += Bc[j] * Fa[j]; }
    stage_state(dv, da);
}

template <int WB, int WG, int WC>
__device__ void pusher_write_obs(const mopa_sawyer_task &T, const WarpWS<WB, WG, WC> &W, float *obs, int lane) {
    // PusherObstacleEnv._get_obs (:185-205): cos / sin of the 4 joint angles, box qpos, joint velocities, box velocity, fingertip xy, goal
    if (lane != 0) return;
    int o = 0;
    for (int k = 0; k < 4; k++) obs[o++] = (float)cos(W.q[T.arm_qadr[k]]);
    for (int k = 0; k < 4; k++) obs[o++] = (float)sin(W.q[T.arm_qadr[k]]);
    for (int k = 0; k < 2; k++) obs[o++] = (float)W.q[T.grip_qadr[k]];
    for (int k = 0; k < 4; k++) obs[o++] = (float)W.v[T.arm_vadr[k]];
    for (int k = 0; k < 2; k++) obs[o++] = (float)W.v[T.grip_vadr[k]];
    double tip[3];
    w_site(tip, W, 0, T.site_grip);
    obs[o++] = (float)tip[0]; obs[o++] = (float)tip[1];
    for (int k = 0; k < 2; k++) obs[o++] = (float)W.q[T.target_qadr[k]];
    while (o < 40) obs[o++] = 0.0f;   // 20 observation floats, row padded to 40
}

template <int WB, int WG, int WC, int ENV_WARPS>
__global__ void __launch_bounds__(ENV_WARPS * 32, 1)
pusher_warp_kernel(int model_slot, const DynDev *__restrict__ mg, mopa_env_buffers B, const float *__restrict__ action, int action_stride,
                   const uint8_t *__restrict__ is_planner, const uint8_t *__restrict__ mask, int n, int fwd, const int32_t *__restrict__ ids,
                   unsigned long long seed, long long env_id_offset, long long *__restrict__ d_episode, const double *__restrict__ qpos0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpWS<WB, WG, WC> &W = reinterpret_cast<WarpWS<WB, WG, WC> *>(smem_raw)[warp];
    const DynDev &m = c_models[model_slot];
    const mopa_sawyer_task &T = c_tasks[model_slot];
    const int t = blockIdx.x * ENV_WARPS + warp;
    const int e = t < n ? (ids ? ids[t] : t) : 0;
    const bool live = t < n && (!mask || mask[e]);
    const int nstage = m.integrator == 0 ? 1 : 4;
    int ncon = 0, nwt = 0;
    double cforce = 0;
    const int4 keep = make_int4(T.body_ee, T.body_cube, T.body_rclaw, T.body_rclaw);
    if (!live) {   // still take part in the CTA barriers of the substep loop
        if (fwd == 0)
            for (int s = 0; s < T.nsub * nstage; s++) w_substep<WB, WG, WC, true>(m, mg, W, 0u, 2, lane, ncon, nwt, cforce, keep, false, true);
        return;
    }
    if (lane < DMAXA) W.ctrl[lane] = 0.0;
    if (lane == 0) W.wn = 0;
    if (lane < WD) W.bias_prev[lane] = 0.0;
    if (lane < m.nd) { int r = lane; while (m.d_parent[r] >= 0) r = m.d_parent[r]; W.blk[lane] = r; }
    __syncwarp();
    if (lane == 0) {   // two kinematic trees with contiguous dofs: [0, k) and [k, nd)
        int k = 1;
        while (k < m.nd && W.blk[k] == W.blk[0]) k++;
        bool two = k < m.nd;
        for (int j = k; j < m.nd && two; j++) two = W.blk[j] == W.blk[k];
        W.blk0 = (two && k == SAWYER_BLK0 && m.nd - k <= WD - SAWYER_BLK0 && !((c_tune.sync_mask >> 9) & 1)) ? k : 0;   // (tuning bit 9: force the dense elimination)
    }
    if (fwd == 3) {   // ---- _reset: rejection sampling
        const unsigned long long gid = (unsigned long long)(env_id_offset + e), ep = (unsigned long long)d_episode[e];
        const int nq = m.nq, nv = m.nv;
        for (int attempt = 0; attempt < 1000; attempt++) {
            const unsigned long long st = gid * 1000003ULL + (unsigned long long)attempt;
            for (int k = lane; k < nq; k += 32) W.q[k] = qpos0[k] + (-0.02 + __dmul_rn(0.04, pz_uniform(seed, st, ep, 4ULL + k)));   // no FMA contraction: the draws are input data, bit-equal to rng.py
            for (int k = lane; k < nv; k += 32) W.v[k] = k >= nv - 4 ? 0.0 : -0.005 + __dmul_rn(0.01, pz_uniform(seed, st, ep, 4ULL + nq + k));
            __syncwarp();
            const double lo[2] = {-0.35, 0.13}, hi[2] = {-0.24, 0.2};
            const double g0 = lo[0] + __dmul_rn(hi[0] - lo[0], pz_uniform(seed, st, ep, 0)), g1 = lo[1] + __dmul_rn(hi[1] - lo[1], pz_uniform(seed, st, ep, 1));
            const double b0 = lo[0] + __dmul_rn(hi[0] - lo[0], pz_uniform(seed, st, ep, 2)), b1 = lo[1] + __dmul_rn(hi[1] - lo[1], pz_uniform(seed, st, ep, 3));
            if (lane == 0) { W.q[nq - 4] = g0; W.q[nq - 3] = g1; W.q[nq - 2] = b0; W.q[nq - 1] = b1; }
            __syncwarp();
            w_substep<WB, WG, WC, true>(m, mg, W, 0u, 3, lane, ncon, nwt, cforce, keep, true, false);
            double d2 = 0;
            for (int k = 0; k < 3; k++) d2 += (W.kxpos[1][k] - W.kxpos[2][k]) * (W.kxpos[1][k] - W.kxpos[2][k]);
            if (ncon == 0 && sqrt(d2) > 0.1 && g0 <= b0) break;
        }
        for (int k = lane; k < nq; k += 32) B.qpos[(size_t)e * nq + k] = W.q[k];
        for (int k = lane; k < nv; k += 32) B.qvel[(size_t)e * nv + k] = W.v[k];
        if (lane < 4 && B.i_term) B.i_term[(size_t)e * 4 + lane] = 0.0;
        pusher_write_obs(T, W, B.obs + (size_t)e * 40, lane);
        if (lane == 0) {
            d_episode[e] += 1;
            B.has_prev[e] = 0; B.ep_len[e] = 0; B.ep_rew[e] = 0.0; B.done[e] = 0; B.success[e] = 0; B.reward[e] = 0.0;
            if (B.unstable) B.unstable[e] = 0;
            if (B.ncon) B.ncon[e] = 0;
        }
        return;
    }
    for (int k = lane; k < m.nq; k += 32) W.q[k] = B.qpos[(size_t)e * m.nq + k];
    for (int k = lane; k < m.nv; k += 32) W.v[k] = B.qvel[(size_t)e * m.nv + k];
    __syncwarp();
    if (fwd == 1) {   // sim.forward() + _get_obs()
        w_substep<WB, WG, WC, true>(m, mg, W, 0u, 0, lane, ncon, nwt, cforce, keep, true, false);
        pusher_write_obs(T, W, B.obs + (size_t)e * 40, lane);
        return;
    }
    const int mode = is_planner ? is_planner[e] : 0;
    const bool planner = mode == 1, had_prev = B.has_prev[e] != 0;
    // desired_state = prev_state + action (the clipped / scaled variants computed before it are dead code in the reference, :253-266)
    double prev = 0, desired = 0, iterm = 0;
    if (lane < 4) {
        prev = (!planner || !had_prev) ? W.q[T.arm_qadr[lane]] : B.prev_state[(size_t)e * 7 + lane];
        desired = prev + (double)action[(size_t)e * action_stride + lane];
        iterm = B.i_term ? B.i_term[(size_t)e * 4 + lane] : 0.0;
    }
    if (mode == 2) w_substep<WB, WG, WC, true>(m, mg, W, 0u, 0, lane, ncon, nwt, cforce, keep, true, false);
    for (int s = 0; s < T.nsub; s++) {
        if (mode != 2 && lane < 4) {   // BaseEnv._get_control: PID on the joint error, re-evaluated before every mj_step
            const double p = T.pid_kp * (desired - W.q[T.arm_qadr[lane]]);
            const double d = T.pid_kd * (0.0 - W.v[T.arm_vadr[lane]]);
            iterm = 0.95 * iterm + T.pid_ki * (prev - W.q[T.arm_qadr[lane]]);
            W.ctrl[lane] = p + d + iterm;
        }
        __syncwarp();
        w_mj_step(m, mg, W, lane, ncon, nwt, cforce, keep, mode != 2, true);
    }
    // instability guard (see env_step_warp_kernel)
    bool unstable = false;
    if (mode != 2) {
        bool bad = false;
        for (int k = lane; k < m.nq; k += 32) { const double x = W.q[k]; if (!(fabs(x) <= 1e10)) bad = true; }
        for (int k = lane; k < m.nv; k += 32) { const double x = W.v[k]; if (!(fabs(x) <= 1e10)) bad = true; }
        unstable = __any_sync(FULL, bad);
        if (unstable) {
            for (int k = lane; k < m.nq; k += 32) W.q[k] = B.qpos[(size_t)e * m.nq + k];
            for (int k = lane; k < m.nv; k += 32) W.v[k] = B.qvel[(size_t)e * m.nv + k];
            __syncwarp();
            w_substep<WB, WG, WC, true>(m, mg, W, 0u, 0, lane, ncon, nwt, cforce, keep, true, false);
            ncon = 0; cforce = 0.0;
        }
    }
    // PusherObstacleEnv.compute_reward (:223-238)
    double reward = 0;
    bool success = false, terminal = false;
    {
        double tip[3], dbg = 0, dbt = 0;
        w_site(tip, W, 0, T.site_grip);
        for (int k = 0; k < 3; k++) {
            dbg += (W.kxpos[1][k] - tip[k]) * (W.kxpos[1][k] - tip[k]);
            dbt += (W.kxpos[1][k] - W.kxpos[2][k]) * (W.kxpos[1][k] - W.kxpos[2][k]);
        }
        dbg = sqrt(dbg); dbt = sqrt(dbt);
        if (dbg < 0.1) reward += 0.1 * (1 - tanh(5 * dbg));
        if (dbt < 0.1) reward += 0.3 * (1 - tanh(5 * dbt));
        if (dbt < T.distance_threshold) { success = true; terminal = true; reward += T.success_reward; }
    }
    if (unstable) { reward = -T.unstable_penalty; success = false; terminal = true; }
    if (mode != 2) pusher_write_obs(T, W, B.obs + (size_t)e * 40, lane);
    __syncwarp();
    // _after_step: joint-limit projection (set_state + forward), episode accounting
    bool clipped = false;
    if (lane < m.nd && m.d_limited[lane] && m.d_qadr[lane] >= 0) {
        double &x = W.q[m.d_qadr[lane]];
        if (x < m.d_range[lane][0]) { x = m.d_range[lane][0]; clipped = true; }
        else if (x > m.d_range[lane][1]) { x = m.d_range[lane][1]; clipped = true; }
    }
    clipped = __any_sync(FULL, clipped) && !unstable;
    __syncwarp();
    if (clipped) w_substep<WB, WG, WC, true>(m, mg, W, 0u, 0, lane, ncon, nwt, cforce, keep, true, false);
    if (!unstable) {
        for (int k = lane; k < m.nq; k += 32) B.qpos[(size_t)e * m.nq + k] = W.q[k];
        for (int k = lane; k < m.nv; k += 32) B.qvel[(size_t)e * m.nv + k] = W.v[k];
        if (mode != 2 && lane < 4) { B.prev_state[(size_t)e * 7 + lane] = desired; if (B.i_term) B.i_term[(size_t)e * 4 + lane] = iterm; }
    }
    if (lane == 0) {
        if (mode != 2) B.has_prev[e] = 1;
        const int len = B.ep_len[e] + 1;
        if (len == T.max_episode_steps) terminal = true;
        B.ep_len[e] = len;
        B.ep_rew[e] += reward;
        B.reward[e] = reward;
        B.done[e] = terminal ? 1 : 0;
        B.success[e] = success ? 1 : 0;
        if (B.ncon) B.ncon[e] = ncon;
        if (B.work && mode != 2) B.work[e] = nwt;
        if (B.cforce) B.cforce[e] = mode != 2 ? cforce : 0.0;
        if (B.unstable) B.unstable[e] = unstable ? 1 : 0;
    }
}

cudaError_t env_tune_set(int prof, int sync_mask) {
    EnvTune t{prof, sync_mask};
    cudaError_t e = cudaMemcpyToSymbol(c_tune, &t, sizeof(t));
    if (e != cudaSuccess) return e;
    unsigned long long z[32] = {0};
    return cudaMemcpyToSymbol(g_prof, z, sizeof(z));
}
cudaError_t env_prof_read(unsigned long long *out) { return cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * 32); }

template <int WB, int WG, int WC, int ENV_WARPS, int ND, int NB>
static cudaError_t launch_env_warp_t(int model_slot, const DynDev *d_model, const mopa_env_buffers &B, const float *action,
                                     int action_stride, const uint8_t *is_planner, const uint8_t *mask, int n, int forward_only,
                                     const int32_t *ids, cudaStream_t stream, int sm_count) {
    static bool attr_set = false;
    static_assert(sizeof(WarpWS<WB, WG, WC>) * ENV_WARPS <= 227 * 1024, "warp workspaces exceed the shared memory of an SM");
    const size_t smem = sizeof(WarpWS<WB, WG, WC>) * ENV_WARPS;
    auto kern = env_step_warp_kernel<WB, WG, WC, ENV_WARPS, ND, NB>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (n <= 0) return cudaSuccess;
    // full CTAs of ENV_WARPS environments (dealing them evenly over rounds x SMs CTAs - fewer warps per SM, the rest idle at the
    // barriers - was measured: 2-5 % slower on all three tasks)
    kern<<<(n + ENV_WARPS - 1) / ENV_WARPS, ENV_WARPS * 32, smem, stream>>>(model_slot, d_model, B, action, action_stride, is_planner, mask, n,
                                                                            forward_only, ids);
    return cudaGetLastError();
}

// The constant bank holds ENV_MODEL_SLOTS scenes; handles are assigned to slots round-robin.  A slot whose tables
// belong to another (older) handle is refreshed before the launch, ordered on the launch stream.
constexpr int ENV_MAX_DEVICES = 16;   // constant banks are per device
static const mopa_env *g_slot_owner[ENV_MAX_DEVICES][ENV_MODEL_SLOTS] = {};
void env_slot_release(const mopa_env *env) {
    for (int d = 0; d < ENV_MAX_DEVICES; d++)
        for (int k = 0; k < ENV_MODEL_SLOTS; k++) if (g_slot_owner[d][k] == env) g_slot_owner[d][k] = nullptr;
}
cudaError_t env_slot_claim(mopa_env *env, cudaStream_t stream, bool force) {
    const int slot = env->model_slot;
    if (slot < 0 || slot >= ENV_MODEL_SLOTS || env->device < 0 || env->device >= ENV_MAX_DEVICES) return cudaErrorInvalidValue;
    if (g_slot_owner[env->device][slot] == env && !force) return cudaSuccess;
    cudaError_t e = cudaMemcpyToSymbolAsync(c_models, &env->h_model, sizeof(DynDev), sizeof(DynDev) * slot, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_tasks, &env->task, sizeof(mopa_sawyer_task), sizeof(mopa_sawyer_task) * slot, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) g_slot_owner[env->device][slot] = env;
    return e;
}

cudaError_t launch_env_warp(mopa_env *env, const mopa_env_buffers &B, const float *action, int action_stride, const uint8_t *is_planner,
                            const uint8_t *mask, int n, int forward_only, const int32_t *ids, cudaStream_t stream) {
    static int small_warps = -1;   // tuning hook: MOPA_ENV_WARPS=7 runs two 7-warp CTAs per SM (smaller barrier domains)
    if (small_warps < 0) { const char *w = getenv("MOPA_ENV_WARPS"); small_warps = (w && atoi(w) == 7) ? 1 : 0; }
    if (env->task.kind == 3)   // PusherObstacle-v0 has its own kernel (RK4 + PID); forward_only maps onto its sim.forward mode
        return launch_pusher(env, B, action, action_stride, is_planner, mask, n, forward_only ? 1 : 0, ids, 0ULL, 0LL, nullptr, stream);
    cudaError_t e = env_slot_claim(env, stream, false);
    if (e != cudaSuccess) return e;
    const int model_slot = env->model_slot, nb = env->h_model.nb, ngeom = env->h_model.ngeom, ngm = env->h_model.ngm;
    const DynDev *d_model = env->d_model;
    const bool small = nb <= 14 && ngeom <= 32 && ngm <= WarpWS<14, 32, 24>::WGM;
#define ENV_LAUNCH(WB_, WG_, WC_, W_, ND_, NB_) launch_env_warp_t<WB_, WG_, WC_, W_, ND_, NB_>(model_slot, d_model, B, action, action_stride, is_planner, mask, n, forward_only, ids, stream, env->sm_count)
    if (env->h_model.nd != 15) return ENV_LAUNCH(DMAXB, DMAXG, 32, 11, 0, 0);   // any other scene: the generic instantiation
    // A launch is `rounds x the latency of one env.step`, and that latency grows with the warps an SM interleaves: a batch that
    // fits one round of 7-warp CTAs (<= 7 environments per SM) runs them instead of filling fewer SMs with 11 / 14 warps.
    const bool one_light_round = n <= 7 * env->sm_count;
    const bool small14 = small && nb == 14;   // the push scene: body count fixed as well
    if (small14 && (small_warps || one_light_round)) return ENV_LAUNCH(14, 32, 24, 7, 15, 14);
    if (small14) return ENV_LAUNCH(14, 32, 24, 14, 15, 14);
    if (nb == 18) return one_light_round ? ENV_LAUNCH(DMAXB, DMAXG, 32, 7, 15, 18) : ENV_LAUNCH(DMAXB, DMAXG, 32, 11, 15, 18);   // lift scene
    if (nb == 19) return one_light_round ? ENV_LAUNCH(DMAXB, DMAXG, 32, 7, 15, 19) : ENV_LAUNCH(DMAXB, DMAXG, 32, 11, 15, 19);   // assembly scene
    if (one_light_round) return ENV_LAUNCH(DMAXB, DMAXG, 32, 7, 15, 0);
    return ENV_LAUNCH(DMAXB, DMAXG, 32, 11, 15, 0);
#undef ENV_LAUNCH
}

cudaError_t launch_pusher(mopa_env *env, const mopa_env_buffers &B, const float *action, int action_stride, const uint8_t *is_planner,
                          const uint8_t *mask, int n, int fwd, const int32_t *ids, unsigned long long seed, long long env_id_offset,
                          long long *d_episode, cudaStream_t stream) {
    cudaError_t e = env_slot_claim(env, stream, false);
    if (e != cudaSuccess) return e;
    constexpr int PW = 8;   // 8 warps per CTA: the Pusher runs with few environments, small CTAs spread them over the SMs
    static bool attr_set = false;
    const size_t smem = sizeof(WarpWS<14, 32, 24>) * PW;
    auto kern = pusher_warp_kernel<14, 32, 24, PW>;
    if (!attr_set) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    kern<<<(n + PW - 1) / PW, PW * 32, smem, stream>>>(env->model_slot, env->d_model, B, action, action_stride, is_planner, mask, n, fwd, ids, seed,
                                                      env_id_offset, d_episode, env->d_qpos0);
    return cudaGetLastError();
}

// 14 workspaces + one 1-warp planner CTA (<= 18 KB + 1 KB reserved each) must fit the 228 KB of an SM together
static_assert(sizeof(WarpWS<14, 32, 24>) * 14 + 1024 + 19 * 1024 <= 228 * 1024, "env CTA leaves no room for a co-resident planner CTA");
size_t env_warp_smem_per_warp() { return sizeof(WarpWS<14, 32, 24>); }

}  // namespace mopa
