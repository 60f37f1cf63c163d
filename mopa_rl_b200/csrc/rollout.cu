// Vectorised MoPA experience collection, device resident: the batched replacement of
// MoPARolloutRunner.run (rl/mopa_rollouts.py:22-399, train branch) and of the planner glue it calls in
// SACAgent (rl/sac_agent.py:145-318: is_planner_ac, convert2planner_displacement, clip_qpos,
// simple_interpolate, plan) and PlannerAgent / SamplingBasedPlanner.plan (re-basing, densification).
//
// One tick = one env.step for every environment.  An environment whose macro action is finished
//   emits its SMDP transition record (ob 40, ac 8, rew, done, intra_steps, env id, ob_next 40),
//   is reset when its episode ended (counter-based draws keyed by env id and episode number),
//   takes the next policy action and either executes it directly (|a| <= omega) or turns it into a plan:
//   displacement map -> target clip -> invalid-target back-off -> straight-line interpolation ->
//   RRT-Connect (asynchronous, on the planner stream; the environment waits) -> densification.
// Every step is a kernel over all environments (or over a compact work list built with atomics); the
// row counts of the state-validity and RRT launches stay on the device, so a tick needs no host
// round trip.  The policy is the only part evaluated outside (between mopa_rollout_pre and _step).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/mopa_b200.h"
#include "env_state.h"
#include "planner_state.h"

void mopa_set_error(const std::string &s);

namespace mopa {
cudaError_t launch_is_valid(const unsigned char *d_blob, const SceneHeader &H, const float *d_qpos, int row_stride, int n,
                            uint32_t *d_out, int exact, int sm_count, cudaStream_t stream, const int *d_n = nullptr, int d_n_mult = 1);

cudaError_t launch_ik_rollout(mopa_env *e, const double *d_qpos, const float *d_policy_ac, const unsigned char *d_need, int body,
                              const double *site_local, const int *joint_dofs, int n_joints, int n, int max_steps, double tol,
                              double action_range, const double *world_lo, const double *world_hi, const double *jlo, const double *jhi,
                              double *d_qpos_scratch, float *d_joint_ac, cudaStream_t stream);

constexpr int RO_JMAX = 16;     // interpolation points checked per straight-line plan
constexpr int RO_NQ = 40;       // qpos capacity (same as the env kernel)
enum { C_MP = 0, C_RL, C_INTERP, C_MP_FAIL, C_APPROX, C_INVALID, C_DENSIFY_FALLBACK, C_EPISODES, C_SUCCESS, C_MP_PATH_LEN,
       C_INTERP_PATH_LEN, C_ENV_STEPS, C_TRANSITIONS, C_RRT_DROPPED, C_RRT_PROBLEMS, C_WAITING, C_REUSED, C_UNSTABLE, C_FB_SIMPLE, C_FB_MAIN, C_COUNT = 24 };

struct RrtBatch {   // one batch of RRT-Connect problems (two of them: being filled / in flight)
    int *cnt;                    // problems queued (may exceed the capacity: clamp)
    int *env;                    // [cap] environment row of each problem
    float *start32, *goal32;     // [cap][row]
    double *start64;             // [cap][nq]  clipped current state (re-basing, passive dims of densified states)
    unsigned long long *keys;    // [cap]
    float *path;                 // [cap][max_path][row]
    int *ids, *plen, *status;    // [cap][max_path], [cap], [cap]
    // densification
    float *dens32;               // [cap][max_path - 1][kmax][row]
    uint32_t *dens_res;          // [cap][max_path - 1][kmax]
    double *wp64;                // [cap][max_path][7] the planner's waypoints re-based on the start state (SamplingBasedPlanner.plan :72-101)
    int *nst;                    // [cap][max_path - 1]  > 0 interior states of the hop, 0 end point only, -(1 + s) fallback path s
    int *ok;                     // [cap]
    // blocked hops -> "simple" planner, then the main planner (SACAgent.simple_interpolate with use_planner, rl/sac_agent.py:300-311)
    int *fb_cnt;                 // hops queued (may exceed the capacity: clamp)
    int *fb_prob, *fb_hop;       // [fb_cap]
    float *fb_start32, *fb_goal32;   // [fb_cap][row]
    unsigned long long *fb_keys; // [fb_cap]
    float *fb_path;              // [fb_cap][fb_max_path][row]
    int *fb_ids, *fb_plen, *fb_status, *fb_status1;   // [fb_cap][fb_max_path], [fb_cap] x 3 (status1: after the simple planner only)
    int *hop_fb;                 // [cap][max_path - 1] fallback slot of a blocked hop, -1 none, -2 dropped (capacity)
};

struct RoDev {   // everything the kernels need, passed by value
    int n, nq, row, max_traj, max_path, kmax, rrt_cap, num_trials, invalid_target_handling, interpolation, task_kind;
    int fb_cap, fb_max_path;     // blocked-hop fallback problems per RRT batch, waypoints per fallback path
    int debug_block_mod;         // test hook, see mopa_rollout_config
    long long *sent;             // [2] records already handed to the replay exchange (ping-pong: read [parity], write [1 - parity])
    int *xfer_overflow;          // mapped host flag: records were overwritten in the ring before they were exchanged
    double omega, action_range, ac_scale, discount, step_size, joint_margin, range;
    unsigned long long seed_env;
    long long env_id_offset;
    double jlo[7], jhi[7], init_qpos[7];
    int arm_qadr[7], target_qadr[2];
    int na;                      // arm joints the policy moves (7 Sawyer, 4 Pusher)
    unsigned unlim;              // bit k: arm joint k is an unlimited hinge (SO(2) in the planner, never clipped): Pusher joint0
    const double *qpos0;         // [nq]
    // per environment
    double *traj;                // [n][max_traj][7]
    int *traj_len, *traj_pos, *executed;
    unsigned char *kind, *pending, *macro_done, *need, *reset_flag, *step_mode, *step_mask;
    float *prev_ob, *ac, *step_action;   // [n][40], [n][8], [n][8]
    double *meta_rew;
    long long *plan_count, *episode_idx;
    long long *macro_index;      // caller-owned [n]
    float *slab;                 // caller-owned [n][92]  records emitted by this tick, dense by environment
    unsigned char *emit_flag;    // caller-owned [n]
    long long *counters;         // caller-owned [C_COUNT]
    float *ring;                 // caller-owned [ring_cap][92]  every record ever emitted (slot = running count % capacity)
    long long ring_cap;
    // planning work lists of the current tick
    int *cnt_plan, *cnt_back;    // device counters
    int *cnt_ids, *ids;          // [2], [n]: environments ordered for the env-step launch (expensive ones first, together)
    int heavy_work;              // Newton steps per env.step above which an environment counts as expensive
    // reuse_data (rl/mopa_rollouts.py:223-302): per-step history of the plan being executed, relabelled records
    int reuse_data, max_reuse;
    int adim, grip_qadr0;        // action entries per environment (7; 8 for the lift task: + gripper), qpos address of rc_close
    double *grip0;               // [n] gripper qpos when the current plan was made (SawyerEnv.form_action with dof == 8, :290-296)
    int ac_normal;               // config.ac_space_type == "normal": displacement = a * action_range (rl/sac_agent.py:160-163, 180-181)
    int discrete;                // config.discrete_action: the policy's ac_type chooses planner / direct execution
    int ik_mode;                 // config.use_ik_target: the action row handed to ro_begin is the IK joint displacement (radians), the
                                 // record keeps the policy's Cartesian action, a planner-branch target is the current state
    float *ctl;                  // [n][8] what is executed: the clipped action, or the IK displacement (+ gripper)
    float *ik_ac;                // [n][8] IK mode: joint displacement rows produced by the IK front end
    double *ik_q;                // [n][nq] IK mode: scratch states of the solver
    unsigned long long seed_reuse;
    float *ob_hist;              // [n][max_traj][40]  observation after step i of the current plan
    double *rew_hist;            // [n][max_traj]      cumulative discounted reward after step i
    unsigned char *done_hist;    // [n][max_traj]
    float *xslab;                // caller-owned [xcap][92] relabelled records of this tick, compact
    double *ep_stats;            // caller-owned [n][5], nullable: episodes finished, sum of ep_len, ep_rew, success, contact force
    double *ep_cforce;           // [n] contact-force sum of the running episode (rl/mopa_rollouts.py:538-539, eval path)
    int *xcount;                 // caller-owned [1]: rows of xslab written by this tick (may exceed xcap: clamp)
    int xcap;
    int *plan_env;               // [n]
    double *tgt64, *c64;         // [n][nq]
    float *q32a;                 // [n][row]          targets
    uint32_t *res_a;             // [n]
    int *back_of_plan;           // [n] back-off slot or -1
    float *q32b;                 // [n][num_trials][row]   back-off candidates (compact by back-off slot)
    uint32_t *res_b;
    float *q32c;                 // [n][RO_JMAX][row] interpolation points
    uint32_t *res_c;
    int *nstep;                  // [n]
    unsigned char *plan_ok;      // [n]
};

__device__ __forceinline__ unsigned long long ro_mix(unsigned long long x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
// mopa_rl_b200/rng.py: uniform01 / normal (counter-based, keyed by seed, stream, counter, dim)
__device__ __forceinline__ double ro_uniform(unsigned long long seed, unsigned long long stream, unsigned long long counter, unsigned long long dim) {
    unsigned long long x = ro_mix(seed ^ (stream * 0x9E3779B97F4A7C15ULL));
    x = ro_mix(x + ((counter << 8) | dim) * 0xD1342543DE82EF95ULL);
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double ro_normal(unsigned long long seed, unsigned long long stream, unsigned long long counter, unsigned long long dim) {
    const double u1 = ro_uniform(seed, stream, counter, dim * 2), u2 = ro_uniform(seed, stream, counter, dim * 2 + 1);
    return sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2);
}
// util/env.py:15-25 joint_convert: unlimited joints are wrapped with period 3.14 (not pi) before they go to the planner
// (SamplingBasedPlanner.convert_nonlimited).  Python's float // and % (CPython float_divmod): the quotient is a whole number here.
__device__ __forceinline__ double ro_joint_convert(double angle) {
    const double w = angle > 0 ? 3.14 : -3.14;
    double mod = fmod(angle, w);
    double div = (angle - mod) / w;
    if (mod != 0.0) { if ((w < 0) != (mod < 0)) { mod += w; div -= 1.0; } }
    else mod = copysign(0.0, w);
    double fl = 0.0;
    if (div != 0.0) { fl = floor(div); if (div - fl > 0.5) fl += 1.0; }
    const bool even = fmod(fl, 2.0) == 0.0;
    return even ? mod : (angle > 0 ? mod - 3.14 : mod + 3.14);
}
__device__ __forceinline__ void ro_count(long long *c, int which, long long v = 1) { atomicAdd((unsigned long long *)(c + which), (unsigned long long)v); }

// ---- 1. finished macro actions: transition records, episode resets (SawyerPushObstacleEnv._reset, :36-51)
__global__ void ro_pre_kernel(RoDev S, mopa_env_buffers B, int nv) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n) return;
    const bool need = S.traj_pos[e] >= S.traj_len[e];
    unsigned char emit = 0, reset = 0;

    if (need) {
        if (S.pending[e]) {
            float *rec = S.slab + (size_t)e * 92;
            for (int k = 0; k < 40; k++) rec[k] = S.prev_ob[(size_t)e * 40 + k];
            for (int k = 0; k < 8; k++) rec[40 + k] = S.ac[(size_t)e * 8 + k];
            rec[48] = (float)S.meta_rew[e];
            rec[49] = S.macro_done[e] ? 1.0f : 0.0f;
            const int ex = S.executed[e] - 1;
            rec[50] = (float)(ex > 0 ? ex : 0);
            rec[51] = (float)(S.env_id_offset + e);
            for (int k = 0; k < 40; k++) rec[52 + k] = B.obs[(size_t)e * 40 + k];
            emit = 1;
            const unsigned long long slot = atomicAdd((unsigned long long *)(S.counters + C_TRANSITIONS), 1ULL) % (unsigned long long)S.ring_cap;
            float *dst = S.ring + slot * 92;
            for (int k = 0; k < 92; k++) dst[k] = rec[k];
            // reuse_data: resample (start, goal) waypoint pairs of the executed plan as extra SMDP transitions
            const int L = S.executed[e];
            if (S.reuse_data && S.kind[e] == 1 && L > 3) {
                const unsigned long long gid = (unsigned long long)(S.env_id_offset + e), mi = (unsigned long long)(S.macro_index[e] - 1);
                const int tries = L < S.max_reuse ? L : S.max_reuse;
                int ps[32], pg[32], np_ = 0;
                const double *tr = S.traj + (size_t)e * S.max_traj * 7;
                const float *oh = S.ob_hist + (size_t)e * S.max_traj * 40;
                const double *rh = S.rew_hist + (size_t)e * S.max_traj;
                for (int t = 0; t < tries; t++) {
                    // np.random.randint(0, L - 1), np.random.randint(start + 1, L) with counter-based draws
                    int start = (int)(ro_uniform(S.seed_reuse, gid, mi, 2ULL * t) * (double)(L - 1));
                    if (start > L - 2) start = L - 2;
                    int goal = start + 1 + (int)(ro_uniform(S.seed_reuse, gid, mi, 2ULL * t + 1) * (double)(L - 1 - start));
                    if (goal > L - 1) goal = L - 1;
                    bool dup = false;
                    for (int k = 0; k < np_; k++) if (ps[k] == start && pg[k] == goal) dup = true;
                    if (dup) continue;
                    ps[np_] = start; pg[np_] = goal; np_++;
                    // env.form_action(traj[goal], traj[start]) -> SACAgent.invert_displacement (piecewise)
                    float ia[7];
                    bool planner_ac = false, valid_ac = true;
                    for (int k = 0; k < S.na; k++) {
                        const double d = tr[goal * 7 + k] - tr[start * 7 + k], ad = fabs(d);
                        const double a = S.ac_normal ? d / S.action_range
                                         : ad < S.ac_scale ? d * (S.omega / S.ac_scale)
                                                           : (d > 0 ? 1.0 : (d < 0 ? -1.0 : 0.0)) *
                                                                 ((ad - S.ac_scale) / ((S.action_range - S.ac_scale) / (1.0 - S.ac_scale)) / ((1.0 - S.ac_scale) / (1.0 - S.omega)) + S.omega);
                        if (a < -S.omega || a > S.omega) planner_ac = true;
                        if (a < -1.0 || a > 1.0) valid_ac = false;
                        ia[k] = (float)a;
                    }
                    if (!planner_ac || !valid_ac) continue;
                    float xr[92];
                    for (int k = 0; k < 40; k++) xr[k] = oh[start * 40 + k];
                    for (int k = 0; k < 8; k++) xr[40 + k] = k < S.na ? ia[k] : 0.0f;
                    xr[47] = S.discrete ? S.ac[(size_t)e * 8 + 7] : 0.0f;   // inter_subgoal_ac["ac_type"] = ac["ac_type"] (:266-267)
                    xr[48] = (float)((rh[goal] - rh[start]) * pow(S.discount, -(double)(start + 1)));
                    xr[49] = S.done_hist[(size_t)e * S.max_traj + goal] ? 1.0f : 0.0f;
                    xr[50] = (float)(goal - start - 1);
                    xr[51] = (float)(S.env_id_offset + e);
                    for (int k = 0; k < 40; k++) xr[52 + k] = oh[goal * 40 + k];
                    if (S.xslab) {   // optional compact per-tick copy of the relabelled records (diagnostics; the exchange reads the ring)
                        const int xs = atomicAdd(S.xcount, 1);
                        if (xs < S.xcap) { float *xo = S.xslab + (size_t)xs * 92; for (int k = 0; k < 92; k++) xo[k] = xr[k]; }
                    }
                    const unsigned long long xslot = atomicAdd((unsigned long long *)(S.counters + C_TRANSITIONS), 1ULL) % (unsigned long long)S.ring_cap;
                    float *xd = S.ring + xslot * 92;
                    for (int k = 0; k < 92; k++) xd[k] = xr[k];
                    ro_count(S.counters, C_REUSED);
                }
            }
            if (S.macro_done[e]) {
                ro_count(S.counters, C_EPISODES);
                if (B.success[e]) ro_count(S.counters, C_SUCCESS);
                if (S.ep_stats) {   // per-environment episode statistics (what run_episode reports: len, rew, episode_success, contact_force)
                    double *st = S.ep_stats + (size_t)e * 5;
                    st[0] += 1.0; st[1] += (double)B.ep_len[e]; st[2] += B.ep_rew[e]; st[3] += B.success[e] ? 1.0 : 0.0; st[4] += S.ep_cforce[e];
                }
                S.ep_cforce[e] = 0.0;
                const unsigned long long gid = (unsigned long long)(S.env_id_offset + e), ep = (unsigned long long)S.episode_idx[e];
                if (S.task_kind != 3) {
                    double *q = B.qpos + (size_t)e * S.nq;
                    for (int k = 0; k < S.nq; k++) q[k] = S.qpos0[k];
                    for (int k = 0; k < S.na; k++) q[S.arm_qadr[k]] = S.init_qpos[k] + 0.02 * ro_normal(S.seed_env, gid, ep, (unsigned long long)k);
                    if (S.task_kind == 0)   // push: the target slides (sawyer_push_obstacle.py:41-47)
                        for (int k = 0; k < 2; k++) q[S.target_qadr[k]] += -0.01 + 0.02 * ro_uniform(S.seed_env, gid, ep, 100ULL + k);
                    for (int k = 0; k < nv; k++) B.qvel[(size_t)e * nv + k] = 0.0;
                    S.episode_idx[e] += 1;
                }   // Pusher: the rejection-sampled reset needs the collision check and runs in the env kernel (mopa_rollout_pre)
                B.ep_len[e] = 0; B.ep_rew[e] = 0.0; B.done[e] = 0; B.success[e] = 0;
                if (B.grasp) B.grasp[e] = 0;
                reset = 1;
            }
        }
        B.has_prev[e] = 0;   // env._reset_prev_state()
    }
    S.need[e] = need ? 1 : 0;
    S.emit_flag[e] = emit;
    S.reset_flag[e] = reset;
    if (e == 0) { *S.cnt_plan = 0; *S.cnt_back = 0; S.cnt_ids[0] = 0; S.cnt_ids[1] = 0; S.counters[C_WAITING] = 0; }
}

// ---- 2. new macro actions: direct action, or planner target (SACAgent.convert2planner_displacement + target clip)
// config.discrete_action (rl/mopa_rollouts.py:86-88, 104-111): the branch is chosen by the policy's ac_type instead of |a| > omega.
// use_ik_target (rec_actions != nullptr): `actions` holds the joint displacement of _cart2dispalcement (radians, not clipped to the action
// box: the reference compares it with omega and divides it by omega as it is), rec_actions the policy's (default[3], quat[4], gripper).
__global__ void ro_begin_kernel(RoDev S, mopa_env_buffers B, const float *__restrict__ actions, const unsigned char *__restrict__ ac_type,
                                const float *__restrict__ rec_actions) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n || !S.need[e]) return;
    float a32[7];
    bool is_mp = false;
    for (int k = 0; k < S.na; k++) {
        float a = actions[(size_t)e * S.adim + k];
        if (!rec_actions) a = a < -1.0f ? -1.0f : (a > 1.0f ? 1.0f : a);
        a32[k] = a;
        S.ctl[(size_t)e * 8 + k] = a;
        S.ac[(size_t)e * 8 + k] = rec_actions ? rec_actions[(size_t)e * 8 + k] : a;
        if (fabs((double)a) > S.omega) is_mp = true;
    }
    if (S.discrete) { is_mp = ac_type[e] != 0; S.ac[(size_t)e * 8 + 7] = is_mp ? 1.0f : 0.0f; }
    if (S.adim == 8) {   // lift: the gripper entry rides along (executed with direct actions and with the last waypoint of a plan)
        float g = actions[(size_t)e * 8 + 7];
        S.ac[(size_t)e * 8 + 7] = g < -1.0f ? -1.0f : (g > 1.0f ? 1.0f : g);
        S.grip0[e] = B.qpos[(size_t)e * S.nq + S.grip_qadr0];
    }
    S.macro_index[e] += 1;
    for (int k = 0; k < 40; k++) S.prev_ob[(size_t)e * 40 + k] = B.obs[(size_t)e * 40 + k];
    S.meta_rew[e] = 0.0; S.executed[e] = 0; S.macro_done[e] = 0; S.pending[e] = 1;
    S.traj_len[e] = 1; S.traj_pos[e] = 0;
    if (!is_mp) { S.kind[e] = 0; ro_count(S.counters, C_RL); return; }
    S.kind[e] = 2;   // failure unless proven otherwise
    const int slot = atomicAdd(S.cnt_plan, 1);
    S.plan_env[slot] = e;
    const double *curr = B.qpos + (size_t)e * S.nq;
    double *tg = S.tgt64 + (size_t)slot * S.nq;
    float *q32 = S.q32a + (size_t)slot * S.row;
    for (int k = 0; k < S.nq; k++) tg[k] = curr[k];
    const double w = S.omega;
    for (int k = 0; k < S.na; k++) {
        const double a = (double)a32[k], aa = fabs(a);
        const double disp = S.ac_normal ? a * S.action_range
                            : aa < w ? a / (w / S.ac_scale)
                                     : (a > 0 ? 1.0 : (a < 0 ? -1.0 : 0.0)) * (S.ac_scale + (S.action_range - S.ac_scale) * ((aa - w) / (1 - w)));
        if (rec_actions) continue;   // use_ik_target: target_qpos stays curr_qpos (rl/mopa_rollouts.py:82, 114-131)
        double t = curr[S.arm_qadr[k]] + disp;
        t = t < S.jlo[k] ? S.jlo[k] : t;
        t = t > S.jhi[k] ? S.jhi[k] : t;
        tg[S.arm_qadr[k]] = t;
    }
    for (int k = 0; k < S.row; k++) q32[k] = k < S.nq ? (float)tg[k] : 0.0f;
}

// ---- 3. invalid targets: candidates of the back-off loop (rl/mopa_rollouts.py:119-143):
//   target += step_size * (curr - target) / |curr - target|   over all nq dims, up to num_trials times.
// One warp per plan slot, lanes over the qpos dims (two per lane); `upto` steps, optionally recording every
// iterate as an fp32 row (the validity queries).  Used by the back-off kernel and by the target choice.
__device__ __forceinline__ void ro_backoff_run(const RoDev &S, const double *c, double &t0, double &t1, int lane, int upto, float *rows) {
    const int k0 = lane, k1 = lane + 32;
    const double c0 = k0 < S.nq ? c[k0] : 0.0, c1 = k1 < S.nq ? c[k1] : 0.0;
    for (int trial = 0; trial < upto; trial++) {
        const double d0 = k0 < S.nq ? c0 - t0 : 0.0, d1 = k1 < S.nq ? c1 - t1 : 0.0;
        double n2 = d0 * d0 + d1 * d1;
        for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        const double nrm = sqrt(n2);
        t0 = t0 + S.step_size * d0 / nrm;
        t1 = t1 + S.step_size * d1 / nrm;
        if (rows) {
            float *q32 = rows + (size_t)trial * S.row;
            if (k0 < S.row) q32[k0] = k0 < S.nq ? (float)t0 : 0.0f;
            if (k1 < S.row) q32[k1] = k1 < S.nq ? (float)t1 : 0.0f;
        }
    }
}
__global__ void ro_backoff_kernel(RoDev S, mopa_env_buffers B) {
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (slot >= *S.cnt_plan) return;
    int bs = -1;
    if (!(S.res_a[slot] & 1u) && S.invalid_target_handling) {
        if (lane == 0) bs = atomicAdd(S.cnt_back, 1);
        bs = __shfl_sync(0xffffffffu, bs, 0);
    }
    if (lane == 0) S.back_of_plan[slot] = bs;
    if (bs < 0) return;
    const int e = S.plan_env[slot];
    const double *c = B.qpos + (size_t)e * S.nq, *tg = S.tgt64 + (size_t)slot * S.nq;
    double t0 = lane < S.nq ? tg[lane] : 0.0, t1 = lane + 32 < S.nq ? tg[lane + 32] : 0.0;
    ro_backoff_run(S, c, t0, t1, lane, S.num_trials, S.q32b + (size_t)bs * S.num_trials * S.row);
}

// ---- 4. target choice, clip_qpos, interpolation points (SACAgent.clip_qpos / simple_interpolate, :237-298).
// One warp per plan slot: the scalar decisions are taken redundantly by every lane, the rows are written by all.
__global__ void ro_interp_kernel(RoDev S, mopa_env_buffers B) {
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (slot >= *S.cnt_plan) return;
    const int e = S.plan_env[slot];
    const double *curr = B.qpos + (size_t)e * S.nq;
    double *tg = S.tgt64 + (size_t)slot * S.nq, *c = S.c64 + (size_t)slot * S.nq;
    bool ok = (S.res_a[slot] & 1u) != 0;
    const int bs = S.back_of_plan[slot];
    if (!ok && bs >= 0) {
        int first = S.num_trials;
        for (int trial = lane; trial < S.num_trials; trial += 32)
            if (S.res_b[(size_t)bs * S.num_trials + trial] & 1u) { first = trial; break; }
        for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
        ok = first < S.num_trials;
        const int upto = ok ? first + 1 : S.num_trials;   // no valid candidate: the last one (and the plan fails)
        double t0 = lane < S.nq ? tg[lane] : 0.0, t1 = lane + 32 < S.nq ? tg[lane + 32] : 0.0;
        ro_backoff_run(S, curr, t0, t1, lane, upto, nullptr);
        if (lane < S.nq) tg[lane] = t0;
        if (lane + 32 < S.nq) tg[lane + 32] = t1;
        __syncwarp();
    }
    if (lane == 0) { S.plan_ok[slot] = ok ? 1 : 0; S.nstep[slot] = 0; }
    if (!ok) { if (lane == 0) { ro_count(S.counters, C_INVALID); ro_count(S.counters, C_MP_FAIL); } return; }
    bool out = false;
    for (int k = 0; k < S.na; k++) { const double x = curr[S.arm_qadr[k]]; if (x < S.jlo[k] || x > S.jhi[k]) out = true; }
    for (int k = lane; k < S.nq; k += 32) c[k] = curr[k];
    __syncwarp();
    if (out && lane < S.na) {
        double x = curr[S.arm_qadr[lane]];
        const double lo = S.jlo[lane] + S.joint_margin, hi = S.jhi[lane] - S.joint_margin;
        x = x < lo ? lo : x;
        x = x > hi ? hi : x;
        c[S.arm_qadr[lane]] = x;
    }
    __syncwarp();
    const double lim = S.ac_scale * 0.8;
    double diff[7], sf = 1.0, run[7];
    for (int k = 0; k < S.na; k++) { diff[k] = tg[S.arm_qadr[k]] - c[S.arm_qadr[k]]; const double s = fabs(diff[k]) / lim; if (s > sf) sf = s; run[k] = c[S.arm_qadr[k]]; }
    int nstep = (int)floor(sf);
    if (nstep > RO_JMAX) nstep = RO_JMAX;
    if (lane == 0) S.nstep[slot] = nstep;
    for (int j = 0; j < RO_JMAX; j++) {
        float *q32 = S.q32c + ((size_t)slot * RO_JMAX + j) * S.row;
        for (int k = lane; k < S.row; k += 32) q32[k] = k < S.nq ? (float)c[k] : 0.0f;
        __syncwarp();
        if (j < nstep) {
            for (int k = 0; k < S.na; k++) run[k] = run[k] + diff[k] / sf;
            if (lane < S.na) {
                double v = run[0];
#pragma unroll
                for (int k = 1; k < 7; k++) if (lane == k) v = run[k];
                q32[S.arm_qadr[lane]] = (float)v;
            }
        }
    }
}

// ---- 5. straight line free -> trajectory; blocked -> RRT-Connect queue (SACAgent.plan, :198-233)
__global__ void ro_interp_finish_kernel(RoDev S, RrtBatch Q) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= *S.cnt_plan || !S.plan_ok[slot]) return;
    const int e = S.plan_env[slot], nstep = S.nstep[slot];
    const double *tg = S.tgt64 + (size_t)slot * S.nq, *c = S.c64 + (size_t)slot * S.nq;
    bool straight = true;
    for (int j = 0; j < nstep; j++) if (!(S.res_c[(size_t)slot * RO_JMAX + j] & 1u)) straight = false;
    if (straight) {
        const double lim = S.ac_scale * 0.8;
        double diff[7], sf = 1.0, run[7];
        for (int k = 0; k < S.na; k++) { diff[k] = tg[S.arm_qadr[k]] - c[S.arm_qadr[k]]; const double s = fabs(diff[k]) / lim; if (s > sf) sf = s; run[k] = c[S.arm_qadr[k]]; }
        double *tr = S.traj + (size_t)e * S.max_traj * 7;
        for (int j = 0; j < nstep; j++)
            for (int k = 0; k < S.na; k++) { run[k] = run[k] + diff[k] / sf; tr[j * 7 + k] = run[k]; }
        for (int k = 0; k < S.na; k++) tr[nstep * 7 + k] = tg[S.arm_qadr[k]];
        S.kind[e] = 1; S.traj_len[e] = nstep + 1; S.traj_pos[e] = 0;
        ro_count(S.counters, C_INTERP);
        ro_count(S.counters, C_INTERP_PATH_LEN, nstep + 1);
        return;
    }
    const int r = atomicAdd(Q.cnt, 1);
    if (r >= S.rrt_cap) { ro_count(S.counters, C_MP_FAIL); ro_count(S.counters, C_RRT_DROPPED); return; }   // kind stays 2
    Q.env[r] = e;
    for (int k = 0; k < S.row; k++) {
        Q.start32[(size_t)r * S.row + k] = k < S.nq ? (float)c[k] : 0.0f;
        Q.goal32[(size_t)r * S.row + k] = k < S.nq ? (float)tg[k] : 0.0f;
    }
    for (int k = 0; k < S.na; k++)
        if ((S.unlim >> k) & 1u) {   // convert_nonlimited on copies of start / goal (sampling_based_planner.py:63-66)
            Q.start32[(size_t)r * S.row + S.arm_qadr[k]] = (float)ro_joint_convert(c[S.arm_qadr[k]]);
            Q.goal32[(size_t)r * S.row + S.arm_qadr[k]] = (float)ro_joint_convert(tg[S.arm_qadr[k]]);
        }
    for (int k = 0; k < S.nq; k++) Q.start64[(size_t)r * S.nq + k] = c[k];
    Q.keys[r] = ((unsigned long long)(S.env_id_offset + e) << 32) + (unsigned long long)S.plan_count[e];   // invariant to batching / GPU count
    S.plan_count[e] += 1;
    S.kind[e] = 3; S.traj_len[e] = 1; S.traj_pos[e] = 0;   // waits for the plan
    ro_count(S.counters, C_RRT_PROBLEMS);
}

// ---- 6a. the planner's waypoints re-based on the (clipped, f64) start state (SamplingBasedPlanner.plan, :72-101).  All joints
// limited (Sawyer): start + (waypoint - first waypoint).  With an unlimited joint (Pusher joint0) the planner saw wrapped
// angles: waypoint deltas are accumulated on the un-wrapped start, going the short way round across +-3.14.
// One thread per problem (the accumulation is sequential; paths have at most max_path rows).
__global__ void ro_rebase_kernel(RoDev S, RrtBatch Q) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int cnt = min(*Q.cnt, S.rrt_cap);
    if (r >= cnt || Q.status[r] != 0) return;
    const int L = Q.plen[r];
    const float *path = Q.path + (size_t)r * S.max_path * S.row;
    const double *st = Q.start64 + (size_t)r * S.nq;
    double *wp = Q.wp64 + (size_t)r * S.max_path * 7;
    for (int k = 0; k < S.na; k++) {
        const int a = S.arm_qadr[k];
        wp[k] = st[a];
        if (S.unlim == 0) {
            const double p0 = (double)path[a];
            for (int i = 1; i < L; i++) wp[(size_t)i * 7 + k] = st[a] + ((double)path[(size_t)i * S.row + a] - p0);
        } else {
            const bool un = (S.unlim >> k) & 1u;
            double acc = st[a];
            for (int i = 1; i < L; i++) {
                const double pv = (double)path[(size_t)(i - 1) * S.row + a], sv = (double)path[(size_t)i * S.row + a];
                double delta = sv - pv;
                if (un && fabs(sv - pv) > 3.14) {
                    if (pv > 0 && sv <= 0) delta = 3.14 - pv + sv + 3.14;
                    else if (pv < 0 && sv > 0) delta = -(3.14 - sv + pv + 3.14);
                }
                acc = acc + delta;
                wp[(size_t)i * 7 + k] = acc;
            }
        }
    }
}
// hop i of problem r: start / end / difference of the re-based waypoints i, i + 1
__device__ __forceinline__ void ro_hop(const RoDev &S, const double *wp, int i, double *hs, double *he, double *diff, double &sf) {
    const double lim = S.ac_scale * 0.8;
    sf = 1.0;
    for (int k = 0; k < S.na; k++) {
        hs[k] = wp[(size_t)i * 7 + k];
        he[k] = wp[(size_t)(i + 1) * 7 + k];
        diff[k] = he[k] - hs[k];
        const double s = fabs(diff[k]) / lim;
        if (s > sf) sf = s;
    }
}

// ---- 6. finished RRT batch: re-base on the start (SamplingBasedPlanner.plan), densify (SACAgent.plan :216-233).
// One warp per problem; lanes stride the hops.
__global__ void ro_rrt_densify_kernel(RoDev S, RrtBatch Q) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int cnt = min(*Q.cnt, S.rrt_cap);
    if (r >= cnt) return;
    const int H = S.max_path - 1;
    // runs on the planner stream: only the batch's own buffers are written here, the environments adopt the result in
    // ro_rrt_finish_kernel (main stream)
    if (Q.status[r] != 0) { if (lane == 0) Q.ok[r] = 0; return; }
    if (lane == 0) Q.ok[r] = 1;
    const int L = Q.plen[r];
    const float *path = Q.path + (size_t)r * S.max_path * S.row;
    const double *st = Q.start64 + (size_t)r * S.nq;
    for (int i = lane; i < H; i += 32) {
        int nst = 0;
        float *rows = Q.dens32 + ((size_t)r * H + i) * S.kmax * S.row;
        if (i < L - 1) {
            double hs[7], he[7], diff[7], sf;
            bool need = false;
            ro_hop(S, Q.wp64 + (size_t)r * S.max_path * 7, i, hs, he, diff, sf);
            for (int k = 0; k < S.na; k++) if (fabs(diff[k]) > S.ac_scale) need = true;
            if (need && S.interpolation) { nst = (int)floor(sf); if (nst > S.kmax) nst = S.kmax; }
            double run[7];
            for (int k = 0; k < S.na; k++) run[k] = hs[k];
            for (int j = 0; j < S.kmax; j++) {
                float *q32 = rows + (size_t)j * S.row;
                for (int k = 0; k < S.row; k++) q32[k] = k < S.nq ? (float)st[k] : 0.0f;
                if (j < nst)
                    for (int k = 0; k < S.na; k++) { run[k] = run[k] + diff[k] / sf; q32[S.arm_qadr[k]] = (float)run[k]; }
            }
        } else {
            for (int j = 0; j < S.kmax; j++)
                for (int k = 0; k < S.row; k++) rows[(size_t)j * S.row + k] = k < S.nq ? (float)st[k] : 0.0f;
        }
        Q.nst[(size_t)r * H + i] = nst;
    }
}
// Hops whose interior states are invalid become planning problems of their own (simple_interpolate with use_planner=True,
// rl/sac_agent.py:300-311): first the "simple" planner (range simple_planner_range, goal_bias is not used by RRT-Connect),
// then the main planner, else the hop keeps only its end point.  One warp per problem, lanes stride the hops.
__global__ void ro_fb_collect_kernel(RoDev S, RrtBatch Q) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int cnt = min(*Q.cnt, S.rrt_cap);
    if (r >= cnt) return;
    const int H = S.max_path - 1, L = Q.ok[r] ? Q.plen[r] : 0;
    const float *path = Q.path + (size_t)r * S.max_path * S.row;
    const double *st = Q.start64 + (size_t)r * S.nq;
    for (int i = lane; i < H; i += 32) {
        int slot = -1;
        if (i < L - 1) {
            const int nst = Q.nst[(size_t)r * H + i];
            bool bad = false;
            for (int j = 0; j < nst; j++) if (!(Q.dens_res[((size_t)r * H + i) * S.kmax + j] & 1u)) bad = true;
            if (S.debug_block_mod > 0 && nst > 0 && (Q.keys[r] + (unsigned long long)i) % (unsigned long long)S.debug_block_mod == 0) bad = true;
            if (bad) {
                slot = atomicAdd(Q.fb_cnt, 1);
                if (slot >= S.fb_cap) slot = -2;
                else {
                    double hs[7], he[7], diff[7], sf;
                    ro_hop(S, Q.wp64 + (size_t)r * S.max_path * 7, i, hs, he, diff, sf);
                    float *s32 = Q.fb_start32 + (size_t)slot * S.row, *g32 = Q.fb_goal32 + (size_t)slot * S.row;
                    for (int k = 0; k < S.row; k++) { const float v = k < S.nq ? (float)st[k] : 0.0f; s32[k] = v; g32[k] = v; }
                    for (int k = 0; k < S.na; k++) {
                        const bool un = (S.unlim >> k) & 1u;   // convert_nonlimited, as for the main problem
                        s32[S.arm_qadr[k]] = (float)(un ? ro_joint_convert(hs[k]) : hs[k]);
                        g32[S.arm_qadr[k]] = (float)(un ? ro_joint_convert(he[k]) : he[k]);
                    }
                    Q.fb_prob[slot] = r; Q.fb_hop[slot] = i;
                    Q.fb_keys[slot] = (Q.keys[r] * 0x9E3779B97F4A7C15ULL) ^ (unsigned long long)(i + 1);
                }
            }
        }
        Q.hop_fb[(size_t)r * H + i] = slot;
    }
}
__global__ void ro_rrt_finish_kernel(RoDev S, RrtBatch Q) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int cnt = min(*Q.cnt, S.rrt_cap);
    if (r >= cnt) return;
    const int e = Q.env[r], H = S.max_path - 1, L = Q.plen[r];
    if (!Q.ok[r]) {   // no exact solution within max_iter (sentinel -4): the plan fails
        if (lane == 0) { S.kind[e] = 2; S.traj_len[e] = 1; S.traj_pos[e] = 0; ro_count(S.counters, C_APPROX); ro_count(S.counters, C_MP_FAIL); }
        return;
    }
    const float *path = Q.path + (size_t)r * S.max_path * S.row;
    const double *st = Q.start64 + (size_t)r * S.nq;
    // pass 1: per hop, what is executed: its interior states + end point, the fallback planner's path, or the end point only
    int total = 0, fallback = 0, fb_simple = 0, fb_main = 0;
    for (int base = 0; base < H; base += 32) {
        const int i = base + lane;
        int c = 0;
        if (i < L - 1) {
            int nst = Q.nst[(size_t)r * H + i];
            const int slot = Q.hop_fb[(size_t)r * H + i];
            if (slot != -1) {   // blocked hop
                if (slot >= 0 && Q.fb_status[slot] == 0 && Q.fb_plen[slot] >= 2) {
                    nst = -(1 + slot);
                    c = Q.fb_plen[slot] - 1;
                    if (Q.fb_status1[slot] == 0) fb_simple++; else fb_main++;
                } else { nst = 0; c = 1; fallback++; }
                Q.nst[(size_t)r * H + i] = nst;
            } else c = nst + 1;
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        total += c;
    }
    for (int o = 16; o > 0; o >>= 1) { fallback += __shfl_xor_sync(0xffffffffu, fallback, o); fb_simple += __shfl_xor_sync(0xffffffffu, fb_simple, o); fb_main += __shfl_xor_sync(0xffffffffu, fb_main, o); }
    __syncwarp();
    if (lane == 0) {
        if (fallback) ro_count(S.counters, C_DENSIFY_FALLBACK, fallback);
        if (fb_simple) ro_count(S.counters, C_FB_SIMPLE, fb_simple);
        if (fb_main) ro_count(S.counters, C_FB_MAIN, fb_main);
    }
    if (total > S.max_traj) {
        if (lane == 0) { S.kind[e] = 2; S.traj_len[e] = 1; S.traj_pos[e] = 0; ro_count(S.counters, C_MP_FAIL); ro_count(S.counters, C_MP); }
        return;
    }
    // pass 2: write the trajectory (offsets by warp scan over chunks of 32 hops)
    double *tr = S.traj + (size_t)e * S.max_traj * 7;
    int carry = 0;
    for (int base = 0; base < H; base += 32) {
        const int i = base + lane;
        const bool live = i < L - 1;
        const int nst = live ? Q.nst[(size_t)r * H + i] : 0;
        const int slot = nst < 0 ? -nst - 1 : -1;
        const int c = live ? (slot >= 0 ? Q.fb_plen[slot] - 1 : nst + 1) : 0;
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int off = carry + incl - c;
        if (live) {
            double hs[7], he[7], diff[7], sf;
            ro_hop(S, Q.wp64 + (size_t)r * S.max_path * 7, i, hs, he, diff, sf);
            if (slot >= 0) {   // the fallback planner's waypoints, re-based on the hop's start (first row dropped: PlannerAgent.plan)
                const float *fp = Q.fb_path + (size_t)slot * S.fb_max_path * S.row;
                for (int k = 0; k < S.na; k++) {
                    const int a = S.arm_qadr[k];
                    if (S.unlim == 0) {
                        for (int j = 1; j <= c; j++) tr[(size_t)(off + j - 1) * 7 + k] = hs[k] + ((double)fp[(size_t)j * S.row + a] - (double)fp[a]);
                    } else {   // accumulate the deltas on the un-wrapped hop start (see ro_rebase_kernel)
                        const bool un = (S.unlim >> k) & 1u;
                        double acc = hs[k];
                        for (int j = 1; j <= c; j++) {
                            const double pv = (double)fp[(size_t)(j - 1) * S.row + a], sv = (double)fp[(size_t)j * S.row + a];
                            double delta = sv - pv;
                            if (un && fabs(sv - pv) > 3.14) {
                                if (pv > 0 && sv <= 0) delta = 3.14 - pv + sv + 3.14;
                                else if (pv < 0 && sv > 0) delta = -(3.14 - sv + pv + 3.14);
                            }
                            acc = acc + delta;
                            tr[(size_t)(off + j - 1) * 7 + k] = acc;
                        }
                    }
                }
            } else {
                double run[7];
                for (int k = 0; k < S.na; k++) run[k] = hs[k];
                for (int j = 0; j < nst; j++)
                    for (int k = 0; k < S.na; k++) { run[k] = run[k] + diff[k] / sf; tr[(size_t)(off + j) * 7 + k] = run[k]; }
                for (int k = 0; k < S.na; k++) tr[(size_t)(off + nst) * 7 + k] = he[k];
            }
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        S.kind[e] = 1; S.traj_len[e] = total; S.traj_pos[e] = 0;
        ro_count(S.counters, C_MP);
        ro_count(S.counters, C_MP_PATH_LEN, total);
    }
}

// ---- replay exchange: the records emitted since the previous call, compact, behind a one-row header (row 0, word 0 = count).
// The local ring is the queue: records [sent, transitions) are pending; at most `cap` leave per call, the rest waits.
__global__ void ro_pack_kernel(RoDev S, float *__restrict__ send, int cap, int parity) {
    const long long total = S.counters[C_TRANSITIONS], sent = S.sent[parity];
    long long pending = total - sent;
    const int k = (int)(pending < (long long)cap ? pending : (long long)cap);
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int i = gt; i < k * 23; i += nt) {   // 92 floats = 23 float4 per record
        const int rec = i / 23, c = i - rec * 23;
        const float4 v = reinterpret_cast<const float4 *>(S.ring + (size_t)((sent + rec) % S.ring_cap) * 92)[c];
        reinterpret_cast<float4 *>(send + (size_t)(1 + rec) * 92)[c] = v;
    }
    if (gt == 0) {
        for (int c = 1; c < 92; c++) send[c] = 0.0f;
        send[0] = __int_as_float(k);
        S.sent[1 - parity] = sent + k;
        if (pending - k > S.ring_cap / 2) *S.xfer_overflow = 1;
    }
}
// Appends the gathered per-rank blocks ([world][1 + cap][92], header row first) to the replicated ring, rank-major: the same
// order on every rank.  One launch; every thread recomputes the (tiny) prefix over the rank counts.
__global__ void replay_append_kernel(const float *__restrict__ recv, int world, int cap, float *__restrict__ ring, long long ring_cap,
                                     long long *__restrict__ size_io, int parity) {
    const long long size0 = size_io[parity];
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    long long off = 0;
    for (int rk = 0; rk < world; rk++) {
        const float *blk = recv + (size_t)rk * (1 + cap) * 92;
        int k = __float_as_int(blk[0]);
        k = k < 0 ? 0 : (k > cap ? cap : k);
        for (int i = gt; i < k * 23; i += nt) {
            const int rec = i / 23, c = i - rec * 23;
            reinterpret_cast<float4 *>(ring + (size_t)((size0 + off + rec) % ring_cap) * 92)[c] =
                reinterpret_cast<const float4 *>(blk + (size_t)(1 + rec) * 92)[c];
        }
        off += k;
    }
    if (gt == 0) size_io[1 - parity] = size0 + off;
}

// ---- 7. stage the action of every environment (direct: ac / omega, plan: env.form_action(next_qpos))
__global__ void ro_stage_kernel(RoDev S, mopa_env_buffers B) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S.n) return;
    const int kind = S.kind[e];
    S.step_mode[e] = (unsigned char)kind;
    S.step_mask[e] = kind != 3;
    if (kind == 3) ro_count(S.counters, C_WAITING);   // environments waiting for their RRT plan in this tick
    // Launch order of the env-step kernel: the warps of a CTA advance in lockstep (stage barriers), so an environment
    // that needs 2-3 Newton steps per substep (arm pushing the cube: ~7 % of them) stalls its 13 neighbours.  Grouping
    // them (cost = Newton steps of their previous env.step) keeps the other CTAs at one step per substep.
    const int heavy = (B.work && B.work[e] > S.heavy_work) ? 1 : 0;
    const int pos = atomicAdd(S.cnt_ids + heavy, 1);
    S.ids[heavy ? pos : S.n - 1 - pos] = e;
    float *sa = S.step_action + (size_t)e * 8;
    if (kind == 0) {
        // direct execution: ac / omega, or the raw action with discrete_action (rl/mopa_rollouts.py:347-352)
        for (int k = 0; k < S.na; k++) sa[k] = S.discrete ? S.ctl[(size_t)e * 8 + k] : (float)((double)S.ctl[(size_t)e * 8 + k] / S.omega);
        if (S.adim == 8) sa[7] = S.ac[(size_t)e * 8 + 7];   // rescaled_ac: only the joint entries are divided by omega (:349-352)
    } else if (kind == 1) {
        int pos = S.traj_pos[e];
        if (pos > S.max_traj - 1) pos = S.max_traj - 1;
        const double *nx = S.traj + ((size_t)e * S.max_traj + pos) * 7;
        for (int k = 0; k < S.na; k++) sa[k] = (float)(nx[k] - B.qpos[(size_t)e * S.nq + S.arm_qadr[k]]);
        // lift: form_action's gripper entry = (gripper qpos of the waypoint = of the plan's start state) - current one; the
        // policy's gripper action replaces it on the last waypoint (rl/mopa_rollouts.py:170-175)
        if (S.adim == 8) sa[7] = (S.traj_pos[e] >= S.traj_len[e] - 1) ? S.ac[(size_t)e * 8 + 7] : (float)(S.grip0[e] - B.qpos[(size_t)e * S.nq + S.grip_qadr0]);
    }
}
// ---- 8. after env.step: discounted macro reward, counters, termination
__global__ void ro_post_kernel(RoDev S, mopa_env_buffers B) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool stepping = e < S.n && S.step_mask[e];
    if (stepping) {
        const int pos = S.traj_pos[e];
        const double disc = S.kind[e] == 1 ? pow(S.discount, (double)pos) : 1.0;
        S.meta_rew[e] += disc * B.reward[e];
        if (S.reuse_data && S.kind[e] == 1 && pos < S.max_traj) {
            float *oh = S.ob_hist + ((size_t)e * S.max_traj + pos) * 40;
            for (int k = 0; k < 40; k++) oh[k] = B.obs[(size_t)e * 40 + k];
            S.rew_hist[(size_t)e * S.max_traj + pos] = S.meta_rew[e];
            S.done_hist[(size_t)e * S.max_traj + pos] = B.done[e];
        }
        S.executed[e] += 1;
        if (B.cforce) S.ep_cforce[e] += B.cforce[e];
        if (B.unstable && B.unstable[e]) ro_count(S.counters, C_UNSTABLE);
        S.traj_pos[e] = pos + 1;
        if (B.done[e]) { S.macro_done[e] = 1; S.traj_len[e] = pos + 1; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, stepping);
    if ((threadIdx.x & 31) == 0 && m) ro_count(S.counters, C_ENV_STEPS, __popc(m));
}

}  // namespace mopa

using namespace mopa;

struct mopa_rollout {
    mopa_env *env = nullptr;
    mopa_planner *planner = nullptr;
    mopa_env_buffers buf;
    RoDev S;
    RrtBatch batch[2];
    int fill = 0;                // batch being filled; the other one may be in flight
    bool inflight = false;
    int max_iter = 1000;
    int simple_max_iter = 25;    // iteration cap of the "simple" planner (stands in for simple_planner_timelimit)
    float simple_range = 0.05f;  // config.simple_planner_range
    int pack_parity = 0;
    mopa_rollout_config cfg;     // the creation-time configuration (the IK front end reads its site / world-box fields)
    int *h_overflow = nullptr;   // mapped host memory: see RoDev::xfer_overflow
    int plan_cta_warps = 1;      // tuning hook: MOPA_PLAN_CTA_WARPS
    cudaStream_t plan_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr, ev_plan0 = nullptr;
    double rrt_last_ms = 0, rrt_sum_ms = 0;   // device time of the finished RRT batches
    long long rrt_batches = 0, tick_launched = 0, rrt_sum_ticks = 0;
    std::vector<void *> allocs;
    long long launches = 0;      // kernels of this library launched so far
    static constexpr int EV_RING = 256;
    cudaEvent_t ev_env0[EV_RING] = {}, ev_env1[EV_RING] = {};   // around the env-step kernel of the latest ticks
    long long ticks = 0;
};

#define RO_TRY(x)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) { mopa_set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); return MOPA_ERR_CUDA; } \
    } while (0)

template <class T>
static cudaError_t ro_alloc(mopa_rollout *r, T **p, size_t count) {
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T));
    if (e != cudaSuccess) return e;
    r->allocs.push_back((void *)*p);
    return cudaMemset(*p, 0, count * sizeof(T));
}

extern "C" {

int mopa_rollout_create(mopa_env *env, mopa_planner *planner, const mopa_env_buffers *buf, const mopa_rollout_config *cfg,
                        int64_t *d_macro_index, float *d_slab, uint8_t *d_emit_flag, float *d_ring, int64_t ring_capacity,
                        int64_t *d_counters, float *d_reuse_slab, int32_t *d_reuse_count, int32_t reuse_capacity, double *d_ep_stats,
                        mopa_rollout **out) {
    if (!env || !planner || !buf || !cfg || !d_macro_index || !d_slab || !d_emit_flag || !d_ring || ring_capacity <= 0 || !d_counters || !out) {
        mopa_set_error("mopa_rollout_create: bad argument");
        return MOPA_ERR_ARG;
    }
    *out = nullptr;
    const DynDev &m = env->h_model;
    if (m.nq > RO_NQ || cfg->n_envs <= 0 || cfg->max_path < 2 || cfg->rrt_capacity <= 0 || planner->scene.hdr.nq != m.nq) {
        mopa_set_error("mopa_rollout_create: configuration outside the compiled limits");
        return MOPA_ERR_ARG;
    }
    {   // limits of the straight-line planner and of the relabelling kernel: refuse, do not clamp
        const double steps = cfg->action_range / (cfg->ac_scale * 0.8);
        if (!(cfg->ac_scale > 0) || steps > (double)RO_JMAX || cfg->max_traj < RO_JMAX + 1) {
            mopa_set_error("mopa_rollout_create: action_range / (0.8 * ac_scale) exceeds the 16 interpolation points of the straight-line planner, or max_traj < 17");
            return MOPA_ERR_ARG;
        }
        if (cfg->reuse_data && (cfg->max_reuse_data < 1 || cfg->max_reuse_data > 32)) {
            mopa_set_error("mopa_rollout_create: max_reuse_data must be in 1..32");
            return MOPA_ERR_ARG;
        }
    }
    if (cfg->use_ik_target && (cfg->discrete_action || env->task.kind != 1 || cfg->ik_body < 0 || cfg->ik_body >= env->h_model.nb || cfg->ik_max_steps < 1)) {
        mopa_set_error("mopa_rollout_create: use_ik_target needs the lift task (8-entry Cartesian actions), no discrete_action and a valid ik_body");
        return MOPA_ERR_ARG;
    }
    mopa_rollout *r = new mopa_rollout();
    r->env = env; r->planner = planner; r->buf = *buf; r->max_iter = cfg->max_iter;
    r->simple_max_iter = cfg->simple_max_iter > 0 ? cfg->simple_max_iter : 0;   // 0: the simple planner gives up at once (tests of the main-planner retry)
    r->simple_range = (float)cfg->simple_planner_range;
    if (const char *w = getenv("MOPA_PLAN_CTA_WARPS")) r->plan_cta_warps = atoi(w);
    RoDev &S = r->S;
    memset(&S, 0, sizeof(S));
    const int n = cfg->n_envs, nq = m.nq, row = planner->scene.hdr.nq4 * 4;
    const double lim = cfg->ac_scale * 0.8;
    S.n = n; S.nq = nq; S.row = row; S.max_traj = cfg->max_traj; S.max_path = cfg->max_path; S.kmax = (int)(cfg->range / lim) + 1;
    S.rrt_cap = cfg->rrt_capacity; S.num_trials = cfg->num_trials; S.invalid_target_handling = cfg->invalid_target_handling;
    S.interpolation = cfg->interpolation;
    S.fb_cap = cfg->rrt_capacity < 256 ? cfg->rrt_capacity : 256; S.fb_max_path = 64;
    S.debug_block_mod = cfg->debug_block_mod;
    S.task_kind = env->task.kind;
    S.omega = cfg->omega; S.action_range = cfg->action_range; S.ac_scale = cfg->ac_scale; S.discount = cfg->discount;
    S.step_size = cfg->step_size; S.joint_margin = cfg->joint_margin; S.range = cfg->range;
    S.seed_env = cfg->seed_env; S.env_id_offset = cfg->env_id_offset;
    S.heavy_work = env->task.nsub + env->task.nsub / 8;
    S.na = env->task.n_arm > 0 ? env->task.n_arm : 7;
    if (S.na > 7) { mopa_set_error("mopa_rollout_create: more than 7 arm joints"); delete r; return MOPA_ERR_ARG; }
    for (int k = 0; k < 7; k++) { S.jlo[k] = cfg->jnt_lo[k]; S.jhi[k] = cfg->jnt_hi[k]; S.init_qpos[k] = cfg->init_qpos[k]; S.arm_qadr[k] = env->task.arm_qadr[k]; }
    for (int k = 0; k < S.na; k++) if (std::isinf(cfg->jnt_lo[k]) || std::isinf(cfg->jnt_hi[k])) S.unlim |= 1u << k;   // unlimited hinge: infinite range
    for (int k = 0; k < 2; k++) S.target_qadr[k] = env->task.target_qadr[k];
    S.macro_index = (long long *)d_macro_index; S.slab = d_slab; S.emit_flag = d_emit_flag; S.counters = (long long *)d_counters;
    S.ring = d_ring; S.ring_cap = ring_capacity;
    S.discrete = cfg->discrete_action ? 1 : 0;
    S.ac_normal = cfg->ac_space_normal ? 1 : 0;
    S.ik_mode = cfg->use_ik_target ? 1 : 0;
    r->cfg = *cfg;
    S.adim = env->task.kind == 1 ? 8 : (env->task.kind == 3 ? 4 : 7);
    S.grip_qadr0 = env->task.grip_qadr[0];
    if (S.adim == 8 && S.discrete) { mopa_set_error("mopa_rollout_create: discrete_action is not built for the 8-D lift action (record slot 47 is taken)"); delete r; return MOPA_ERR_ARG; }
    S.reuse_data = cfg->reuse_data ? 1 : 0;
    S.max_reuse = cfg->max_reuse_data < 1 ? 1 : cfg->max_reuse_data;
    S.ep_stats = d_ep_stats;
    S.seed_reuse = cfg->seed_reuse;
    if (d_reuse_slab && d_reuse_count && reuse_capacity > 0) { S.xslab = d_reuse_slab; S.xcount = d_reuse_count; S.xcap = reuse_capacity; }
    cudaError_t e = cudaSetDevice(env->device);
#define A(ptr, count) if (e == cudaSuccess) e = ro_alloc(r, &ptr, (size_t)(count))
    double *qpos0 = nullptr;
    A(qpos0, nq);
    if (e == cudaSuccess) e = cudaMemcpy(qpos0, cfg->qpos0, sizeof(double) * nq, cudaMemcpyHostToDevice);
    S.qpos0 = qpos0;
    A(S.traj, (size_t)n * S.max_traj * 7);
    A(S.traj_len, n); A(S.traj_pos, n); A(S.executed, n);
    A(S.kind, n); A(S.pending, n); A(S.macro_done, n); A(S.need, n); A(S.reset_flag, n); A(S.step_mode, n); A(S.step_mask, n);
    A(S.prev_ob, (size_t)n * 40); A(S.ac, (size_t)n * 8); A(S.ctl, (size_t)n * 8); A(S.step_action, (size_t)n * 8);
    if (S.ik_mode) { A(S.ik_ac, (size_t)n * 8); A(S.ik_q, (size_t)n * nq); }
    A(S.meta_rew, n); A(S.plan_count, n); A(S.episode_idx, n);
    A(S.grip0, n); A(S.cnt_plan, 1); A(S.cnt_back, 1); A(S.cnt_ids, 2); A(S.ids, n); A(S.ep_cforce, n);
    if (S.reuse_data) { A(S.ob_hist, (size_t)n * S.max_traj * 40); A(S.rew_hist, (size_t)n * S.max_traj); A(S.done_hist, (size_t)n * S.max_traj); }
    A(S.plan_env, n); A(S.tgt64, (size_t)n * nq); A(S.c64, (size_t)n * nq); A(S.q32a, (size_t)n * row); A(S.res_a, n);
    A(S.back_of_plan, n); A(S.q32b, (size_t)n * S.num_trials * row); A(S.res_b, (size_t)n * S.num_trials);
    A(S.q32c, (size_t)n * RO_JMAX * row); A(S.res_c, (size_t)n * RO_JMAX); A(S.nstep, n); A(S.plan_ok, n);
    const size_t cap = S.rrt_cap, H = S.max_path - 1;
    for (int b = 0; b < 2; b++) {
        RrtBatch &Q = r->batch[b];
        A(Q.cnt, 1); A(Q.env, cap); A(Q.start32, cap * row); A(Q.goal32, cap * row); A(Q.start64, cap * nq); A(Q.keys, cap);
        A(Q.path, cap * S.max_path * row); A(Q.ids, cap * S.max_path); A(Q.plen, cap); A(Q.status, cap);
        A(Q.dens32, cap * H * S.kmax * row); A(Q.dens_res, cap * H * S.kmax); A(Q.nst, cap * H); A(Q.ok, cap);
        A(Q.wp64, cap * S.max_path * 7);
        const size_t fc = S.fb_cap, fp = S.fb_max_path;
        A(Q.fb_cnt, 1); A(Q.fb_prob, fc); A(Q.fb_hop, fc); A(Q.fb_start32, fc * row); A(Q.fb_goal32, fc * row); A(Q.fb_keys, fc);
        A(Q.fb_path, fc * fp * row); A(Q.fb_ids, fc * fp); A(Q.fb_plen, fc); A(Q.fb_status, fc); A(Q.fb_status1, fc); A(Q.hop_fb, cap * H);
    }
    A(S.sent, 2);
#undef A
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&r->h_overflow, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) { *r->h_overflow = 0; e = cudaHostGetDevicePointer((void **)&S.xfer_overflow, r->h_overflow, 0); }
    if (e == cudaSuccess) {   // the planner stream outranks the env-step kernel: its small CTAs take the first SM slots that free up
        int lo = 0, hi = 0;
        e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&r->plan_stream, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&r->ev_done);
    if (e == cudaSuccess) e = cudaEventCreate(&r->ev_plan0);
    for (int k = 0; k < mopa_rollout::EV_RING && e == cudaSuccess; k++) { e = cudaEventCreate(&r->ev_env0[k]); if (e == cudaSuccess) e = cudaEventCreate(&r->ev_env1[k]); }
    // episode counters start at 1: episode 0 was drawn by the initial reset of the environments
    if (e == cudaSuccess) {
        std::vector<long long> one(n, 1);
        e = cudaMemcpy(S.episode_idx, one.data(), sizeof(long long) * n, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        mopa_set_error(std::string("mopa_rollout_create: ") + cudaGetErrorString(e));
        mopa_rollout_destroy(r);
        return MOPA_ERR_CUDA;
    }
    *out = r;
    return MOPA_OK;
}

void mopa_rollout_destroy(mopa_rollout *r) {
    if (!r) return;
    cudaSetDevice(r->env->device);
    cudaDeviceSynchronize();
    for (void *p : r->allocs) cudaFree(p);
    if (r->h_overflow) cudaFreeHost(r->h_overflow);
    if (r->plan_stream) cudaStreamDestroy(r->plan_stream);
    if (r->ev_ready) cudaEventDestroy(r->ev_ready);
    if (r->ev_done) cudaEventDestroy(r->ev_done);
    if (r->ev_plan0) cudaEventDestroy(r->ev_plan0);
    for (int k = 0; k < mopa_rollout::EV_RING; k++) { if (r->ev_env0[k]) cudaEventDestroy(r->ev_env0[k]); if (r->ev_env1[k]) cudaEventDestroy(r->ev_env1[k]); }
    delete r;
}

static int ro_finalize_rrt(mopa_rollout *r, cudaStream_t st) {
    RoDev &S = r->S;
    RrtBatch &Q = r->batch[1 - r->fill];
    const int warps_blocks = (S.rrt_cap * 32 + 127) / 128;
    RO_TRY(cudaStreamWaitEvent(st, r->ev_done, 0));
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r->ev_plan0, r->ev_done) == cudaSuccess) { r->rrt_last_ms = ms; r->rrt_sum_ms += ms; r->rrt_batches += 1; r->rrt_sum_ticks += r->ticks - r->tick_launched; }
    }
    ro_rrt_finish_kernel<<<warps_blocks, 128, 0, st>>>(S, Q);
    RO_TRY(cudaGetLastError());
    r->inflight = false;
    r->launches += 1;
    return MOPA_OK;
}

/* The whole planning pipeline of a batch, on the planner stream (under the env-step kernels of the following ticks):
 * RRT-Connect -> densification states -> their validity -> blocked hops -> simple planner -> main planner on what is left. */
static int ro_launch_rrt(mopa_rollout *r, RrtBatch &Q) {
    RoDev &S = r->S;
    mopa_planner *p = r->planner;
    cudaStream_t ps = r->plan_stream;
    const int warps_blocks = (S.rrt_cap * 32 + 127) / 128, H = S.max_path - 1;
    RO_TRY(launch_plan(p, Q.start32, Q.goal32, S.row, Q.keys, S.rrt_cap, r->max_iter, Q.path, Q.ids, S.max_path, Q.plen, Q.status, nullptr, nullptr,
                       ps, Q.cnt, r->plan_cta_warps));   // 1-warp CTAs: small enough to share an SM with an env-step CTA
    ro_rebase_kernel<<<(S.rrt_cap + 127) / 128, 128, 0, ps>>>(S, Q);
    ro_rrt_densify_kernel<<<warps_blocks, 128, 0, ps>>>(S, Q);
    RO_TRY(launch_is_valid(p->d_blob, p->scene.hdr, Q.dens32, S.row, S.rrt_cap * H * S.kmax, Q.dens_res, 0, p->sm_count, ps, Q.cnt, H * S.kmax));
    RO_TRY(cudaMemsetAsync(Q.fb_cnt, 0, sizeof(int), ps));
    ro_fb_collect_kernel<<<warps_blocks, 128, 0, ps>>>(S, Q);
    RO_TRY(cudaGetLastError());
    RO_TRY(launch_plan(p, Q.fb_start32, Q.fb_goal32, S.row, Q.fb_keys, S.fb_cap, r->simple_max_iter, Q.fb_path, Q.fb_ids, S.fb_max_path, Q.fb_plen,
                       Q.fb_status, nullptr, nullptr, ps, Q.fb_cnt, r->plan_cta_warps, r->simple_range, 0, 0ULL));
    RO_TRY(cudaMemcpyAsync(Q.fb_status1, Q.fb_status, sizeof(int) * S.fb_cap, cudaMemcpyDeviceToDevice, ps));
    RO_TRY(launch_plan(p, Q.fb_start32, Q.fb_goal32, S.row, Q.fb_keys, S.fb_cap, r->max_iter, Q.fb_path, Q.fb_ids, S.fb_max_path, Q.fb_plen,
                       Q.fb_status, nullptr, nullptr, ps, Q.fb_cnt, r->plan_cta_warps, 0.f, 1, 0xA5A5A5A5A5A5A5A5ULL));
    r->launches += 7;
    return MOPA_OK;
}

/* First half of a tick: finished RRT batch -> trajectories; finished macro actions -> transition records
 * (slab / emit_flag) and episode resets (+ sim.forward of the reset environments). */
int mopa_rollout_pre(mopa_rollout *r, int32_t wait_rrt, void *stream) {
    if (!r) return MOPA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    RoDev &S = r->S;
    RO_TRY(cudaSetDevice(r->env->device));
    // Keep the host at most one tick ahead of the device: the decision to finalise the RRT batch is taken on the host
    // (event query), so a host that has queued many ticks would see the planner's progress that many ticks late.
    if (r->ticks >= 2) RO_TRY(cudaEventSynchronize(r->ev_env1[(r->ticks - 2) % mopa_rollout::EV_RING]));
    if (r->inflight) {
        cudaError_t q = wait_rrt ? cudaEventSynchronize(r->ev_done) : cudaEventQuery(r->ev_done);
        if (q == cudaSuccess) { int rc = ro_finalize_rrt(r, st); if (rc) return rc; }
        else if (q != cudaErrorNotReady) RO_TRY(q);
    }
    const int blocks = (S.n + 127) / 128;
    if (S.xslab) RO_TRY(cudaMemsetAsync(S.xcount, 0, sizeof(int), st));
    ro_pre_kernel<<<blocks, 128, 0, st>>>(S, r->buf, r->env->h_model.nv);
    RO_TRY(cudaGetLastError());
    if (S.task_kind == 3)   // PusherObstacleEnv._reset: rejection sampling with the collision check, draws keyed by (env id, episode)
        RO_TRY(launch_pusher(r->env, r->buf, nullptr, 0, nullptr, S.reset_flag, S.n, 3, nullptr, S.seed_env, S.env_id_offset, S.episode_idx, st));
    else
        RO_TRY(launch_env_warp(r->env, r->buf, nullptr, 0, nullptr, S.reset_flag, S.n, 1, nullptr, st));
    r->launches += 2;
    return MOPA_OK;
}

/* Second half: d_actions [n][7] = the policy's output for every environment (used where a new macro
 * action starts).  Planning glue, RRT launch, env.step for every non-waiting environment, bookkeeping. */
int mopa_rollout_step(mopa_rollout *r, const float *d_actions, void *stream) {
    if (r && r->S.ik_mode) { mopa_set_error("mopa_rollout_step: the handle runs use_ik_target, call mopa_rollout_step_ik"); return MOPA_ERR_ARG; }
    return mopa_rollout_step_discrete(r, d_actions, nullptr, stream);
}
int mopa_rollout_step_ik(mopa_rollout *r, const float *d_actions, void *stream) {
    if (!r || !r->S.ik_mode) { mopa_set_error("mopa_rollout_step_ik: the handle was not created with use_ik_target"); return MOPA_ERR_ARG; }
    return mopa_rollout_step_discrete(r, d_actions, nullptr, stream);
}

/* Same with the policy's ac_type [n] (0 = direct execution, 1 = motion planner); required when the handle was created
 * with discrete_action (scripts/3d/push/mopa_discrete.sh), ignored otherwise. */
int mopa_rollout_step_discrete(mopa_rollout *r, const float *d_actions, const uint8_t *d_ac_type, void *stream) {
    if (!r || !d_actions) return MOPA_ERR_ARG;
    if (r->S.discrete && !d_ac_type) { mopa_set_error("mopa_rollout_step: the handle runs discrete_action, ac_type is required"); return MOPA_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    RoDev &S = r->S;
    mopa_planner *p = r->planner;
    RO_TRY(cudaSetDevice(r->env->device));
    const int blocks = (S.n + 127) / 128;
    RrtBatch &Q = r->batch[r->fill];
    const float *begin_actions = d_actions, *rec_actions = nullptr;
    if (S.ik_mode) {   // _cart2dispalcement for the environments that start a macro action: Cartesian action -> joint displacement row
        const mopa_rollout_config &C = r->cfg;
        int dofs[7];
        for (int k = 0; k < 7; k++) dofs[k] = r->env->task.arm_dof[k];
        RO_TRY(launch_ik_rollout(r->env, r->buf.qpos, d_actions, S.need, C.ik_body, C.ik_site_local, dofs, S.na, S.n, C.ik_max_steps, C.ik_tol,
                                 S.action_range, C.ik_world_lo, C.ik_world_hi, S.jlo, S.jhi, S.ik_q, S.ik_ac, st));
        begin_actions = S.ik_ac; rec_actions = d_actions;
        r->launches += 1;
    }
    ro_begin_kernel<<<blocks, 128, 0, st>>>(S, r->buf, begin_actions, d_ac_type, rec_actions);
    RO_TRY(launch_is_valid(p->d_blob, p->scene.hdr, S.q32a, S.row, S.n, S.res_a, 0, p->sm_count, st, S.cnt_plan, 1));
    ro_backoff_kernel<<<(S.n * 32 + 127) / 128, 128, 0, st>>>(S, r->buf);
    RO_TRY(launch_is_valid(p->d_blob, p->scene.hdr, S.q32b, S.row, S.n * S.num_trials, S.res_b, 0, p->sm_count, st, S.cnt_back, S.num_trials));
    ro_interp_kernel<<<(S.n * 32 + 127) / 128, 128, 0, st>>>(S, r->buf);
    RO_TRY(launch_is_valid(p->d_blob, p->scene.hdr, S.q32c, S.row, S.n * RO_JMAX, S.res_c, 0, p->sm_count, st, S.cnt_plan, RO_JMAX));
    ro_interp_finish_kernel<<<blocks, 128, 0, st>>>(S, Q);
    RO_TRY(cudaGetLastError());
    if (!r->inflight) {
        // hand the filled batch to the planner stream (it runs under the env-step kernels); start filling the other one
        RO_TRY(cudaEventRecord(r->ev_ready, st));
        RO_TRY(cudaStreamWaitEvent(r->plan_stream, r->ev_ready, 0));
        RO_TRY(cudaEventRecord(r->ev_plan0, r->plan_stream));
        r->tick_launched = r->ticks;
        { const int rc = ro_launch_rrt(r, Q); if (rc) return rc; }
        RO_TRY(cudaEventRecord(r->ev_done, r->plan_stream));
        r->inflight = true;
        r->fill = 1 - r->fill;
        RO_TRY(cudaMemsetAsync(r->batch[r->fill].cnt, 0, sizeof(int), st));
    }
    ro_stage_kernel<<<blocks, 128, 0, st>>>(S, r->buf);
    RO_TRY(cudaGetLastError());
    const int ev = (int)(r->ticks % mopa_rollout::EV_RING);
    RO_TRY(cudaEventRecord(r->ev_env0[ev], st));
    RO_TRY(launch_env_warp(r->env, r->buf, S.step_action, 8, S.step_mode, S.step_mask, S.n, 0, S.ids, st));
    RO_TRY(cudaEventRecord(r->ev_env1[ev], st));
    ro_post_kernel<<<blocks, 128, 0, st>>>(S, r->buf);
    RO_TRY(cudaGetLastError());
    r->launches += 10;   // begin, 3 x validity, back-off, interpolation, interpolation finish, stage, env step, post
    r->ticks += 1;
    return MOPA_OK;
}

/* Replay exchange, step 1 (see include/mopa_b200.h). */
int mopa_rollout_pack(mopa_rollout *r, float *d_send, int32_t cap, void *stream) {
    if (!r || !d_send || cap <= 0) { mopa_set_error("mopa_rollout_pack: bad argument"); return MOPA_ERR_ARG; }
    if (*r->h_overflow) {
        mopa_set_error("mopa_rollout_pack: transition records were overwritten before they were exchanged (the slab capacity is too small for the emission rate)");
        return MOPA_ERR_OVERFLOW;
    }
    RO_TRY(cudaSetDevice(r->env->device));
    const int blocks = (cap * 23 + 255) / 256 < 296 ? (cap * 23 + 255) / 256 : 296;
    ro_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(r->S, d_send, cap, r->pack_parity);
    RO_TRY(cudaGetLastError());
    r->pack_parity ^= 1;
    r->launches += 1;
    return MOPA_OK;
}

/* Replay exchange, step 2: gathered blocks -> replicated ring. */
int mopa_replay_append(const float *d_recv, int32_t world, int32_t cap, float *d_ring, int64_t ring_capacity, int64_t *d_size2, int32_t parity,
                       void *stream) {
    if (!d_recv || world <= 0 || cap <= 0 || !d_ring || ring_capacity <= 0 || !d_size2 || (parity != 0 && parity != 1)) {
        mopa_set_error("mopa_replay_append: bad argument");
        return MOPA_ERR_ARG;
    }
    const int blocks = (cap * 23 + 255) / 256 < 296 ? (cap * 23 + 255) / 256 : 296;
    replay_append_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_recv, world, cap, d_ring, (long long)ring_capacity, (long long *)d_size2, parity);
    RO_TRY(cudaGetLastError());
    return MOPA_OK;
}

/* 1 while an RRT batch is in flight or problems are queued behind it (host view; the queue length lives on the device). */
int mopa_rollout_busy(mopa_rollout *r) { return r && r->inflight ? 1 : 0; }

/* Diagnostics of the asynchronous planner: out[0] = device ms of the last finished RRT batch, out[1] = batches finished,
 * out[2] = mean device ms per batch, out[3] = mean ticks between launch and finalisation. */
int mopa_rollout_rrt_stats(mopa_rollout *r, double *out4) {
    if (!r || !out4) return MOPA_ERR_ARG;
    out4[0] = r->rrt_last_ms; out4[1] = (double)r->rrt_batches;
    out4[2] = r->rrt_batches ? r->rrt_sum_ms / r->rrt_batches : 0.0;
    out4[3] = r->rrt_batches ? (double)r->rrt_sum_ticks / r->rrt_batches : 0.0;
    return MOPA_OK;
}

/* Kernels of this library launched so far by the handle. */
int64_t mopa_rollout_launches(mopa_rollout *r) { return r ? r->launches : 0; }

/* Mean device time (ms) of the env-step kernel over the latest n_last ticks (synchronises the device). */
int mopa_rollout_env_ms(mopa_rollout *r, int32_t n_last, double *out_ms) {
    if (!r || !out_ms || n_last <= 0) return MOPA_ERR_ARG;
    RO_TRY(cudaSetDevice(r->env->device));
    RO_TRY(cudaDeviceSynchronize());
    long long have = r->ticks < mopa_rollout::EV_RING ? r->ticks : mopa_rollout::EV_RING;
    if (n_last > have) n_last = (int32_t)have;
    double sum = 0;
    for (int k = 0; k < n_last; k++) {
        const int ev = (int)((r->ticks - 1 - k) % mopa_rollout::EV_RING);
        float ms = 0;
        RO_TRY(cudaEventElapsedTime(&ms, r->ev_env0[ev], r->ev_env1[ev]));
        sum += ms;
    }
    *out_ms = n_last ? sum / n_last : 0.0;
    return MOPA_OK;
}

}  // extern "C"
