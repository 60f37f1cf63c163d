/* C ABI of libmopa_b200.so — the B200-native drop-in for the native half of MoPA-RL's
 * experience-collection path.
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * clvrai/mopa-rl checkout).  Conventions: every function returns 0 on success and a negative
 * code on failure (message via mopa_last_error()); no exception crosses the ABI; the caller
 * owns every buffer; handles are owned by the library; one host thread per handle; device
 * entry points enqueue on the caller-supplied cudaStream_t (passed as void*) and do not
 * synchronise; *_host entry points take host memory, perform the copies themselves and
 * return when the result is in the caller's buffer.
 */
#ifndef MOPA_B200_H
#define MOPA_B200_H
#include <stdint.h>

#include "mopa_dyn_desc.h"
#include "mopa_model_desc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mopa_planner mopa_planner;

#define MOPA_OK 0
#define MOPA_ERR_ARG (-1)
#define MOPA_ERR_CUDA (-2)
#define MOPA_ERR_MODEL (-3)
#define MOPA_ERR_OVERFLOW (-6)   /* a fixed-capacity queue lost data (see mopa_rollout_pack) */

/* mopa_is_valid_* flags */
#define MOPA_VALID_FAST 0        /* result word: bit0 = valid; other bits 0 */
#define MOPA_VALID_FIRST_PAIR 1  /* no early-out; bits 8.. = 1 + canonical index of the first offending pair */

/* planner status codes written by mopa_plan_* (the reference signals them with sentinel rows:
 * KinematicPlanner.cpp:181-184 (-5, invalid goal) and :249-250 (-4, no exact solution)) */
#define MOPA_PLAN_OK 0
#define MOPA_PLAN_NOT_EXACT (-4)
#define MOPA_PLAN_INVALID_GOAL (-5)

/* Last error message of the calling thread ("" if none). */
const char *mopa_last_error(void);

/* Library / device probe: returns the number of visible CUDA devices (<0 on error). */
int mopa_device_count(void);

/* Replaces KinematicPlanner::KinematicPlanner (motion_planners/KinematicPlanner.cpp:42-120,
 * bound by PyKinematicPlanner.__cinit__, motion_planners/planner.pyx:33-41).
 *   model              compiled scene (instead of the XML file name; MuJoCo is not available)
 *   passive_qpos_idx   qpos indices frozen during a plan ("passive_joint_idx")
 *   ignored_pairs      ordered geom-id pairs whose contacts never invalidate ("ignored_contacts")
 *   contact_threshold  a contact invalidates the state iff dist <= contact_threshold
 *   range              RRT-Connect extension range (setRange, KinematicPlanner.cpp:101)
 *   resolution         state validity checking resolution (0.005 in KinematicPlanner.cpp:87)
 *   seed               ompl::RNG::setSeed(seed) (KinematicPlanner.cpp:93); keys the counter-based RNG
 */
int mopa_planner_create(const mopa_model_desc *model, const int32_t *passive_qpos_idx, int32_t n_passive,
                        const int32_t *ignored_pairs, int32_t n_ignored, double contact_threshold, double range,
                        double resolution, uint64_t seed, int32_t device, mopa_planner **out);
void mopa_planner_destroy(mopa_planner *p);

/* Scene facts: qpos length, candidate-pair count, number of active (planned) joints. */
int mopa_planner_info(const mopa_planner *p, int32_t *nq, int32_t *n_pairs, int32_t *n_active);
/* Canonical candidate pair list (mjModel geom ids), n_pairs entries each. */
int mopa_planner_pairs(const mopa_planner *p, int32_t *geom1, int32_t *geom2);
/* Host-only view of the pair table the kernels sweep (no device needed): which canonical candidate pairs are kept after the
 * build-time reach analysis (pairs whose bounding spheres cannot touch for any joint configuration are dropped, see
 * csrc/scene_build.cu).  kept[i] = 1 / 0 per canonical pair (n_pairs entries, may be NULL);
 * stats[0..7] = table entries incl. padding, kept pairs, cull runs, dropped pairs, canonical pairs, table bytes, frame floats per state, 0. */
int mopa_scene_pair_table(const mopa_model_desc *model, const int32_t *ignored_pairs, int32_t n_ignored, double contact_threshold,
                          int32_t *stats, uint8_t *kept, int32_t n_pairs);

/* Replaces MujocoStateValidityChecker::isValid (mujoco_ompl_interface.cpp:909-978), batched.
 *   d_qpos      device, n rows of fp32 qpos, row_stride floats apart (row_stride >= nq)
 *   d_result    device, n result words (see MOPA_VALID_*)
 */
int mopa_is_valid_batch(mopa_planner *p, const float *d_qpos, int32_t row_stride, int32_t n, uint32_t *d_result,
                        int32_t flags, void *stream);

/* Replaces KinematicPlanner::isValidState (KinematicPlanner.cpp:253-286, planner.pyx:51-52) for
 * n host states of nq doubles each.  valid[i] in {0,1}; words (nullable) receives the result words. */
int mopa_is_valid_host(mopa_planner *p, const double *qpos, int32_t n, uint8_t *valid, uint32_t *words, int32_t flags);

/* Same check for n host rows that are already fp32 (row_stride floats apart, pinned memory
 * recommended): the rows are streamed to the device in chunks on two streams so that the
 * host->device copy, the kernel and the device->host copy of the result words overlap. */
int mopa_is_valid_host_f32(mopa_planner *p, const float *qpos, int32_t row_stride, int32_t n, uint32_t *words, int32_t flags);

/* The reference's own calling convention (KinematicPlanner::isValidState(std::vector<double> state_vec),
 * KinematicPlanner.cpp:253-286; SamplingBasedPlanner.isValidState in sampling_based_planner.py): a state is the vector of
 * the ACTIVE joints (the planner's state space, n_active values, in the order of mopa_planner_info / the non-passive qpos
 * indices), the passive joints keep the values of base_qpos (nq floats, host) - the planner's own mjData in the reference.
 * n host states of n_active floats each (pinned memory recommended); 5x less host->device traffic than full qpos rows for
 * the Sawyer scenes.  Rows are expanded on the device and checked by the same kernel: identical result words. */
int mopa_is_valid_active_host_f32(mopa_planner *p, const float *active, int32_t n, const float *base_qpos, uint32_t *words, int32_t flags);

/* Node capacity of each RRT tree (default 4096).  A tree that fills up stops growing. */
int mopa_planner_set_max_nodes(mopa_planner *p, int32_t max_nodes);

/* Replaces KinematicPlanner::plan (KinematicPlanner.cpp:125-251, planner.pyx:45-46), batched: n
 * independent problems, one warp each.  RRT-Connect with `range`, edge validation at
 * `resolution`, passive dims frozen at the start values (KinematicPlanner.cpp:166,238).
 *   d_start, d_goal  device, n rows of fp32 qpos (row_stride floats apart)
 *   d_keys           device, n 64-bit problem keys for the counter-based RNG (with `seed`)
 *   max_iter         cap on main-loop iterations (stands in for the wall-clock `timelimit`)
 *   d_path           device, [n][max_path][row_stride] fp32 waypoints incl. the start row
 *   d_node_ids       device, [n][max_path] tree-node index of each waypoint (bit 30: goal tree)
 *   d_path_len       device, [n] rows written (0 on failure)
 *   d_status         device, [n] MOPA_PLAN_*
 *   d_iters,d_nodes  device, nullable: [n] iterations used, [n][2] tree sizes
 */
int mopa_plan_batch(mopa_planner *p, const float *d_start, const float *d_goal, int32_t row_stride, const uint64_t *d_keys,
                    int32_t n, int32_t max_iter, float *d_path, int32_t *d_node_ids, int32_t max_path, int32_t *d_path_len,
                    int32_t *d_status, int32_t *d_iters, int32_t *d_nodes, void *stream);

/* Host-buffer variant (what PyKinematicPlanner.plan hands over: vectors of nq doubles).
 *   path [n][max_path][nq] doubles, node_ids [n][max_path] (nullable), path_len/status/iters [n]. */
int mopa_plan_host(mopa_planner *p, const double *start, const double *goal, const uint64_t *keys, int32_t n, int32_t max_iter,
                   double *path, int32_t *node_ids, int32_t max_path, int32_t *path_len, int32_t *status, int32_t *iters);

/* ------------------------------------------------------------------------------------------
 * Vectorised environments (replaces BaseEnv.step / SawyerPushObstacleEnv._step, compute_reward,
 * _get_obs and BaseEnv._after_step: env/base.py:232-314, env/sawyer/sawyer_push_obstacle.py:71-208).
 * One thread integrates one environment; the 75 mj_step-equivalent substeps of an env.step stay
 * on chip.  All state lives in caller-owned device arrays (torch tensors on the Python side).
 */
typedef struct mopa_env mopa_env;

typedef struct mopa_sawyer_task {
    int32_t kind;                 /* 0: SawyerPushObstacle-v0, 1: SawyerLiftObstacle-v0 (8-D action: 7 joints + gripper; body_cube = can),
                                   * 2: SawyerAssemblyObstacle-v0 (body_cube = peg, body_rclaw = the part that
                                   * carries the hole sites, site_right_eef / site_left_eef = pegHead / pegEnd),
                                   * 3: PusherObstacle-v0 (env/pusher/pusher_obstacle.py; BASELINE configs[0]): n_arm = 4 hinge joints
                                   * (arm_* entries 0..3), body_ee = fingertip, site_grip = site "fingertip", body_cube = box,
                                   * body_rclaw = target, grip_qadr / grip_vadr = the box slides (qpos[-2:]), target_qadr = the
                                   * goal slides (qpos[-4:-2]); velocity actuators driven by the PID law of BaseEnv._get_control
                                   * (env/base.py:200-209) re-evaluated before each of the nsub mj_steps (RK4, dt 0.01) */
    int32_t arm_qadr[7], arm_vadr[7], arm_dof[7];   /* qpos / qvel addresses and simulated-dof indices of right_j0..6 */
    int32_t grip_qadr[2], grip_vadr[2];             /* rc_close, lc_close */
    int32_t body_ee, body_cube, body_rclaw, body_lclaw; /* simulated-body indices */
    int32_t target_qadr[2];
    int32_t max_episode_steps, nsub;                /* 250; int(frame_dt / timestep) = 75 */
    double site_right_eef[3], site_left_eef[3], site_grip[3];  /* site positions in their body frames */
    double target_base[3];                          /* body_pos of the target body */
    double ac_scale, distance_threshold, success_reward;
    double site_hole[3], site_hole_bottom[3];       /* assembly: sites "hole" / "hole_bottom" in their body frame */
    /* lift (env/sawyer/sawyer_lift_obstacle.py:92-148): simulated-geom indices of the can and of the left / right finger geoms
     * (l_finger_g0, l_finger_g1, l_fingertip_g0 / r_*; -1 = absent) whose contacts define has_grasp, and z of body bin1 */
    int32_t geom_cube, geom_lfinger[3], geom_rfinger[3];
    int32_t n_arm;                /* arm joints the policy moves: 7 (0 is read as 7), 4 for the Pusher */
    double bin_z;
    double unstable_penalty;      /* env_config["unstable_penalty"] (env/base.py:36, default 0): subtracted from the reward of a step whose
                                   * simulation diverged (BaseEnv._do_simulation / _after_step, env/base.py:300-304, 388-400) */
    double pid_kp, pid_kd, pid_ki;  /* Pusher: gains of BaseEnv._get_control (150, 20, 0.1; integral leak 0.95) */
} mopa_sawyer_task;

typedef struct mopa_env_buffers {   /* device pointers, n_envs rows each */
    double *qpos;        /* [n][nq] */
    double *qvel;        /* [n][nv] */
    double *prev_state;  /* [n][7]  SawyerEnv._prev_state */
    double *bias_prev;   /* [n][16] qfrc_bias of the previous mj_step on the simulated dofs */
    uint8_t *has_prev;   /* [n]     _prev_state is not None */
    int32_t *ep_len;     /* [n]     _episode_length */
    double *ep_rew;      /* [n]     _episode_reward */
    float *obs;          /* [n][40] observation in the reference's key order */
    double *reward;      /* [n] */
    uint8_t *done;       /* [n]     _terminal after _after_step */
    uint8_t *success;    /* [n]     _success */
    int32_t *ncon;       /* [n]     contacts in the last substep */
    int32_t *work;       /* [n]     nullable: Newton steps spent by the latest env.step (cost feedback: callers group
                          *         expensive environments into the same CTAs, see mopa_rollout_step) */
    double *cforce;      /* [n]     nullable: BaseEnv.get_contact_force() after the step (env/base.py:568-581): sum over the
                          *         contacts of the last substep of |f_normal| + |f_tangent1| + |f_tangent2| */
    uint8_t *grasp;      /* [n]     nullable (lift): bit 0 / 1 = the can touched a left / right finger geom in the contact list of the
                          *         latest simulated step; a planner-failure step (no mj_step) evaluates compute_reward on this stale
                          *         list exactly as the reference reads mjData.contact there (rl/mopa_rollouts.py:312) */
    double *i_term;      /* [n][4]  nullable (Pusher): integral term of the PID law, carried across env.steps (BaseEnv._i_term) */
    uint8_t *unstable;   /* [n]     nullable: 1 when the latest step produced a non-finite / huge (> 1e10, mjMAXVAL) state: the step is
                          *         discarded (state row untouched), the episode terminates with -unstable_penalty, and the caller
                          *         resets the environment (BaseEnv._do_simulation: reset() + _fail, env/base.py:388-400) */
} mopa_env_buffers;

/* sizeof() of the ABI structs as this library was compiled (bindings compare them with their own layout at load time):
 * out[0..4] = mopa_model_desc, mopa_dyn_desc, mopa_sawyer_task, mopa_env_buffers, mopa_rollout_config. */
int mopa_abi_sizes(int32_t *out5);

int mopa_env_create(const mopa_dyn_desc *dyn, const mopa_sawyer_task *task, int32_t device, mopa_env **out);
void mopa_env_destroy(mopa_env *e);
int mopa_env_enable_contacts(mopa_env *e, int32_t on);
/* diagnostics: per-stage clock sums of the env-step kernel, collected when MOPA_ENV_PROF=1 is set at create time */
int mopa_env_debug_prof(mopa_env *e, uint64_t *out32);
/* sim.forward() + _get_obs() for the listed envs (d_ids nullable = all n): refreshes bias_prev and obs
 * from qpos/qvel.  Called after reset / set_state. */
int mopa_env_forward(mopa_env *e, const mopa_env_buffers *buf, const int32_t *d_ids, int32_t n, void *stream);
/* env.step(action, is_planner) for every env whose mask byte is non-zero (d_mask nullable = all).
 * d_action [n][action_stride] fp32 (first 7 used); d_is_planner [n] (nullable = all 0):
 *   0  direct action (scaled by ac_scale), 1  planner waypoint (is_planner=True: joint displacement),
 *   2  planner failure: compute_reward + _after_step without simulation (rl/mopa_rollouts.py:304-327). */
int mopa_env_step(mopa_env *e, const mopa_env_buffers *buf, const float *d_action, int32_t action_stride,
                  const uint8_t *d_is_planner, const uint8_t *d_mask, int32_t n_envs, void *stream);

/* PusherObstacleEnv._reset (env/pusher/pusher_obstacle.py:40-67) for every env whose mask byte is non-zero (d_mask nullable = all):
 * rejection sampling of (goal, box, joint noise, velocity noise) until the state has no contact, the box is farther than 0.1 from
 * the target and goal_x <= box_x.  Draws come from the counter-based generator keyed by (seed, global env id * 1000003 + attempt,
 * episode); d_episode int64[n] holds each env's episode number and is incremented for the envs that were reset.  qpos0: host,
 * nq doubles (the keyframe the noise is added to).  Also clears the episode counters and refreshes the observation. */
int mopa_env_reset_pusher(mopa_env *e, const mopa_env_buffers *buf, const uint8_t *d_mask, uint64_t seed, int64_t env_id_offset,
                          int64_t *d_episode, const double *qpos0, int32_t n_envs, void *stream);

/* Batched inverse kinematics on a site pose - replaces qpos_from_site_pose (env/inverse_kinematics.py:18-135, called by
 * MoPARolloutRunner._cart2dispalcement, rl/mopa_rollouts.py:683-728): damped least squares with the reference's constants
 * (regularisation 3e-2, update-norm cap 2.0, progress threshold 20, rot_weight 1).
 *   d_qpos [n][nq] start states; d_target_pos [n][3]; d_target_quat [n][4] wxyz, nullable (position only)
 *   body / site_local: simulated-body index and local position of the site; joint_dofs: simulated-dof indices of the
 *   movable joints (env.robot_joints); outputs: d_qpos_out [n][nq], d_err [n] final error norm, d_steps [n], d_success [n] */
int mopa_ik_batch(mopa_env *e, const double *d_qpos, const double *d_target_pos, const double *d_target_quat, int32_t body,
                  const double *site_local, const int32_t *joint_dofs, int32_t n_joints, int32_t n, int32_t max_steps, double tol,
                  double *d_qpos_out, double *d_err, int32_t *d_steps, uint8_t *d_success, void *stream);

/* ------------------------------------------------------------------------------------------
 * Device-resident experience collection (replaces MoPARolloutRunner.run, rl/mopa_rollouts.py:22-399, and the
 * planner glue of SACAgent / PlannerAgent / SamplingBasedPlanner.plan it calls, rl/sac_agent.py:145-318).
 * One tick = mopa_rollout_pre, the caller's policy on the observations of all environments, mopa_rollout_step.
 */
typedef struct mopa_rollout mopa_rollout;
typedef struct mopa_rollout_config {
    int32_t n_envs, max_iter, max_path, max_traj, rrt_capacity, num_trials, invalid_target_handling, interpolation;
    double omega, action_range, ac_scale, discount, step_size, joint_margin, range;
    uint64_t seed_env;            /* seed of the reset draws (keyed by env id and episode number) */
    int64_t env_id_offset;        /* global id of environment row 0 (shards of a multi-GPU run) */
    double jnt_lo[7], jnt_hi[7];  /* joint ranges of the arm */
    double init_qpos[7];          /* arm pose the reset noise is added to */
    const double *qpos0;          /* host, nq: reset pose of everything else */
    int32_t reuse_data, max_reuse_data;   /* rl/mopa_rollouts.py:223-302: up to max_reuse_data (<= 32) relabelled records per executed plan */
    uint64_t seed_reuse;          /* seed of the (start, goal) draws, keyed by env id and macro-action index */
    int32_t discrete_action;      /* config.discrete_action (rl/mopa_rollouts.py:86-88): ac_type picks planner / direct execution;
                                     direct actions are not divided by omega; record slot 47 carries ac_type */
    int32_t ac_space_normal;      /* config.ac_space_type == "normal" (scripts/3d/{lift,assembly}/mopa_discrete.sh): planner displacement =
                                     a * action_range and relabelled action = d / action_range (rl/sac_agent.py:160-163, 180-181); 0 = piecewise */
    /* SACAgent._simple_planner (rl/sac_agent.py:98-110, used by simple_interpolate(use_planner=True), :300-311): a densification hop
     * whose interior is blocked is re-planned with RRT-Connect at `simple_planner_range` for `simple_max_iter` iterations (the cap that
     * stands in for simple_planner_timelimit), then with the main planner (`range`, `max_iter`), else only its end point is kept. */
    double simple_planner_range;
    int32_t simple_max_iter;
    int32_t debug_block_mod;      /* test hook, 0 = off: treat the interior of densification hop i of a problem as blocked when
                                     (problem key + i) % debug_block_mod == 0, so that conformance tests reach the fallback planners
                                     (naturally ~0.3 % of the RRT plans have such a hop) */
    /* config.use_ik_target (the MoPA-SAC IK presets, rl/mopa_rollouts.py:90-101, 355-362, 683-728): the policy acts in Cartesian
     * space - action row = (default[3], quat[4], gripper) - and every macro action starts with _cart2dispalcement: target position
     * = clip(site + action_range * default, world box), target orientation = mat2quat(site_xmat)[[3, 0, 1, 1]] (x) quat / |quat|
     * (the reference's index list, kept), qpos_from_site_pose(max_steps, tol) on the arm joints, joint displacement = clipped
     * result - current.  The displacement then takes the place of the action: |d_k| > omega -> planner branch, whose target is
     * the CURRENT state in the reference (the increment sits behind `if not config.use_ik_target`, :114-131: the plan is a
     * two-step hold), else env.step(d / omega (+ gripper)).  Use mopa_rollout_step_ik. */
    int32_t use_ik_target;
    int32_t ik_body;              /* simulated-body index that carries the site config.ik_target */
    double ik_site_local[3];      /* the site's position in that body's frame (its local rotation must be the identity) */
    double ik_world_lo[3], ik_world_hi[3];   /* env.min_world_size / env.max_world_size */
    int32_t ik_max_steps;         /* 100 in the reference */
    int32_t ik_pad_;
    double ik_tol;                /* 1e-2 in the reference */
} mopa_rollout_config;
/* Counter slots of d_counters (int64[24]; the named ones below, the rest reserved). */
#define MOPA_RO_COUNTERS "mp,rl,interpolation,mp_fail,approximate,invalid,densify_fallback,episodes,success,mp_path_len,interpolation_path_len,env_steps,transitions,rrt_dropped,rrt_problems,waiting,reused,unstable,fb_simple,fb_main"
/* Caller-owned device buffers: d_macro_index int64[n] (policy calls per env), d_slab float[n][92] + d_emit_flag
 * uint8[n] (records emitted by the latest tick, dense by environment), d_ring float[ring_capacity][92] (all
 * records, slot = running count % capacity), d_counters int64[24]; with reuse_data: d_reuse_slab float[reuse_capacity][92] +
 * d_reuse_count int32[1] (relabelled records of the latest tick, compact; rows beyond the capacity only reach the ring);
 * d_ep_stats double[n][5], nullable: per environment the number of finished episodes and the sums of their length, reward,
 * success flag and contact force (the quantities MoPARolloutRunner.run_episode reports, rl/mopa_rollouts.py:401-681). */
int mopa_rollout_create(mopa_env *env, mopa_planner *planner, const mopa_env_buffers *buf, const mopa_rollout_config *cfg,
                        int64_t *d_macro_index, float *d_slab, uint8_t *d_emit_flag, float *d_ring, int64_t ring_capacity,
                        int64_t *d_counters, float *d_reuse_slab, int32_t *d_reuse_count, int32_t reuse_capacity, double *d_ep_stats,
                        mopa_rollout **out);
void mopa_rollout_destroy(mopa_rollout *r);
/* wait_rrt != 0: block until the RRT batch in flight (if any) is done, then finalise it. */
int mopa_rollout_pre(mopa_rollout *r, int32_t wait_rrt, void *stream);
/* d_actions: device float[n][7] (float[n][8] for the lift task: 7 joint entries + gripper), the policy's action for every environment. */
int mopa_rollout_step(mopa_rollout *r, const float *d_actions, void *stream);
/* discrete_action handles: d_ac_type uint8[n] (0 = direct execution, 1 = motion planner), the policy's ac["ac_type"]. */
int mopa_rollout_step_discrete(mopa_rollout *r, const float *d_actions, const uint8_t *d_ac_type, void *stream);
/* use_ik_target handles: d_actions float[n][8] = (default[3], quat[4], gripper) per environment; the IK solve runs on the device
 * for the environments that start a macro action, the records keep the policy's Cartesian action. */
int mopa_rollout_step_ik(mopa_rollout *r, const float *d_actions, void *stream);
/* Replicated replay (the consumer rl/dataset.py:7-37 replaces): step 1 of the per-tick exchange.  Copies the records emitted
 * since the previous call - at most `cap`, the rest stays queued in the ring - to d_send float[1 + cap][92]: row 0 is a header whose
 * first word holds the record count (int32 bits), rows 1.. the records.  The caller all-gathers d_send across ranks (NCCL) and hands
 * the result to mopa_replay_append.  Returns MOPA_ERR_OVERFLOW when queued records were overwritten in the ring before they left. */
int mopa_rollout_pack(mopa_rollout *r, float *d_send, int32_t cap, void *stream);
/* Step 2: d_recv float[world][1 + cap][92] (the gathered blocks; world = 1: the send buffer itself) -> rows appended to the replicated
 * ring d_ring float[ring_capacity][92] in rank-major order (identical on every rank).  d_size2 int64[2]: running record count, read
 * from [parity] and written to [1 - parity] (the caller alternates parity = 0, 1, 0, ... so that one launch suffices). */
int mopa_replay_append(const float *d_recv, int32_t world, int32_t cap, float *d_ring, int64_t ring_capacity, int64_t *d_size2, int32_t parity,
                       void *stream);
int mopa_rollout_busy(mopa_rollout *r);
/* Diagnostics of the asynchronous planner: out4 = {device ms of the last finished RRT batch, batches finished, mean device
 * ms per batch, mean ticks between launch and finalisation}. */
int mopa_rollout_rrt_stats(mopa_rollout *r, double *out4);
/* Kernels of this library launched so far through the handle. */
int64_t mopa_rollout_launches(mopa_rollout *r);
/* Mean device time (ms) of the env-step kernel over the latest n_last ticks (synchronises the device). */
int mopa_rollout_env_ms(mopa_rollout *r, int32_t n_last, double *out_ms);

#ifdef __cplusplus
}
#endif
#endif
