// Vectorised Sawyer environments: env.step on sm_100a, one thread per environment.
//
// Replaces BaseEnv.step -> SawyerPushObstacleEnv._step -> 75 x sim.step() -> compute_reward /
// _get_obs -> BaseEnv._after_step (env/base.py:232-314, env/sawyer/sawyer_push_obstacle.py:71-208,
// env/sawyer/sawyer.py:317-338).  The physics is dyn.cuh; this file adds the reference's env
// logic around it: the latched _prev_state, action clipping, gravity compensation by the
// previous step's qfrc_bias, reward / success, the 40-float observation and the joint-limit
// projection + episode accounting of _after_step.  Per env.step HBM sees the state row in and
// out plus the observation; the substeps never leave the SM.
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

#include "../../include/mopa_b200.h"
#include "dyn.cuh"
#include "contact.cuh"

#include "env_state.h"


namespace mopa {

static void fill_model(const mopa_dyn_desc *d, DynDev &m) {
    memset(&m, 0, sizeof(m));
    m.nq = d->nq; m.nv = d->nv; m.nb = d->nb; m.nd = d->nd; m.nact = d->nact; m.ngeom = d->ngeom; m.npair = d->npair;
    m.iterations = d->iterations; m.h = d->timestep; m.tolerance = d->tolerance;
    m.integrator = d->integrator;
    for (int k = 0; k < 3; k++) m.g[k] = d->gravity[k];
    for (int i = 0; i < d->nb; i++) {
        m.b_parent[i] = d->b_parent[i]; m.b_jtype[i] = d->b_jtype[i]; m.b_qadr[i] = d->b_qadr[i]; m.b_vadr[i] = d->b_vadr[i];
        m.b_dadr[i] = d->b_dadr[i]; m.b_qpos0[i] = d->b_qpos0[i]; m.b_mass[i] = d->b_mass[i];
        for (int k = 0; k < 3; k++) {
            m.b_pos[i][k] = d->b_pos[3 * i + k]; m.b_rootpos[i][k] = d->b_rootpos[3 * i + k]; m.b_jaxis[i][k] = d->b_jaxis[3 * i + k];
            m.b_jpos[i][k] = d->b_jpos[3 * i + k]; m.b_ipos[i][k] = d->b_ipos[3 * i + k]; m.b_inertia[i][k] = d->b_inertia[3 * i + k];
        }
        for (int k = 0; k < 4; k++) { m.b_quat[i][k] = d->b_quat[4 * i + k]; m.b_rootquat[i][k] = d->b_rootquat[4 * i + k]; m.b_iquat[i][k] = d->b_iquat[4 * i + k]; }
    }
    for (int i = 0; i < d->nd; i++) {
        m.d_body[i] = d->d_body[i]; m.d_qadr[i] = d->d_qadr[i]; m.d_vadr[i] = d->d_vadr[i]; m.d_limited[i] = d->d_limited[i];
        m.d_armature[i] = d->d_armature[i]; m.d_damping[i] = d->d_damping[i]; m.d_margin[i] = d->d_margin[i];
        for (int k = 0; k < 2; k++) { m.d_range[i][k] = d->d_range[2 * i + k]; m.d_solref[i][k] = d->d_solref[2 * i + k]; }
        for (int k = 0; k < 5; k++) m.d_solimp[i][k] = d->d_solimp[5 * i + k];
        const int b = m.d_body[i];
        if (i > 0 && m.d_body[i - 1] == b) { m.d_parent[i] = i - 1; continue; }
        int p = m.b_parent[b];
        while (p >= 0 && m.b_jtype[p] < 0) p = m.b_parent[p];
        m.d_parent[i] = p < 0 ? -1 : m.b_dadr[p] + (m.b_jtype[p] == 0 ? 5 : 0);
    }
    for (int i = 0; i < d->nact; i++) {
        m.a_dof[i] = d->a_dof[i]; m.a_kind[i] = d->a_kind[i]; m.a_ctrllimited[i] = d->a_ctrllimited[i]; m.a_forcelimited[i] = d->a_forcelimited[i];
        m.a_kp[i] = d->a_kp[i]; m.a_kv[i] = d->a_kv[i]; m.a_gear[i] = d->a_gear[i];
        for (int k = 0; k < 2; k++) { m.a_ctrlrange[i][k] = d->a_ctrlrange[2 * i + k]; m.a_forcerange[i][k] = d->a_forcerange[2 * i + k]; }
    }
    for (int i = 0; i < d->ngeom; i++) {
        m.g_body[i] = d->g_body[i]; m.g_type[i] = d->g_type[i]; m.g_rbound[i] = d->g_rbound[i]; m.g_margin[i] = d->g_margin[i];
        for (int k = 0; k < 3; k++) { m.g_pos[i][k] = d->g_pos[3 * i + k]; m.g_size[i][k] = d->g_size[3 * i + k]; m.g_friction[i][k] = d->g_friction[3 * i + k]; }
        for (int k = 0; k < 4; k++) m.g_quat[i][k] = d->g_quat[4 * i + k];
        for (int k = 0; k < 2; k++) m.g_solref[i][k] = d->g_solref[2 * i + k];
        for (int k = 0; k < 5; k++) m.g_solimp[i][k] = d->g_solimp[5 * i + k];
    }
    for (int i = 0; i < d->npair; i++) { m.p_g1[i] = d->p_g1[i]; m.p_g2[i] = d->p_g2[i]; }
    // geom rotations: world matrix of a static geom / local matrix of a moving one; moving geoms whose local
    // rotation is not the identity get a slot in the kernel's per-substep world-matrix cache
    m.ngm = 0;
    for (int i = 0; i < d->ngeom; i++) {
        d_q2m(m.g_mat[i], m.g_quat[i]);
        const bool ident = m.g_quat[i][0] == 1.0 && m.g_quat[i][1] == 0.0 && m.g_quat[i][2] == 0.0 && m.g_quat[i][3] == 0.0;
        if (m.g_body[i] < 0) m.g_mslot[i] = -2;
        else if (ident) m.g_mslot[i] = -1;
        else if (m.ngm < DMAXGM) { m.gm_geom[m.ngm] = i; m.g_mslot[i] = m.ngm++; }
        else m.ngm = DMAXGM + 1;   // too many: rejected by mopa_env_create
    }
    m.enable_contacts = 1;
}

}  // namespace mopa

void mopa_set_error(const std::string &s);
#define ENV_TRY(x)                                                                                  \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) { mopa_set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); return MOPA_ERR_CUDA; } \
    } while (0)

extern "C" {

int mopa_env_create(const mopa_dyn_desc *dyn, const mopa_sawyer_task *task, int32_t device, mopa_env **out) {
    if (!dyn || !task || !out) { mopa_set_error("mopa_env_create: bad argument"); return MOPA_ERR_ARG; }
    *out = nullptr;
    if (dyn->nb > mopa::DMAXB || dyn->nd > mopa::DMAXD || dyn->nact > mopa::DMAXA || dyn->ngeom > mopa::DMAXG || dyn->npair > 640 ||
        dyn->nq > 36 || dyn->nv > 36) {
        mopa_set_error("mopa_env_create: scene exceeds the compiled limits of the env kernel");
        return MOPA_ERR_MODEL;
    }
    if (task->kind < 0 || task->kind > 3) { mopa_set_error("mopa_env_create: task kind must be 0 (push), 1 (lift), 2 (assembly) or 3 (pusher)"); return MOPA_ERR_MODEL; }
    if (task->kind == 3 && (task->n_arm != 4 || dyn->nb > 14 || dyn->ngeom > 32)) { mopa_set_error("mopa_env_create: the pusher task needs n_arm = 4 and a scene within the small workspace"); return MOPA_ERR_MODEL; }
    if (task->kind != 3 && dyn->integrator != 0) { mopa_set_error("mopa_env_create: the Sawyer tasks integrate with Euler (RK4 is built for the pusher task)"); return MOPA_ERR_MODEL; }
    if (task->kind == 1 && (task->geom_cube < 0 || task->geom_cube >= dyn->ngeom)) { mopa_set_error("mopa_env_create: lift task without the can's contact geom"); return MOPA_ERR_MODEL; }
    mopa_env *e = new mopa_env();
    e->device = device;
    e->task = *task;
    mopa::fill_model(dyn, e->h_model);
    if (e->h_model.ngm > mopa::DMAXGM) { mopa_set_error("mopa_env_create: more rotated moving geoms than the env kernel caches"); delete e; return MOPA_ERR_MODEL; }
    cudaError_t err = cudaSetDevice(device);
    if (err == cudaSuccess) err = cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (err == cudaSuccess) err = cudaMalloc(&e->d_model, sizeof(mopa::DynDev));
    if (err == cudaSuccess) err = cudaMemcpy(e->d_model, &e->h_model, sizeof(mopa::DynDev), cudaMemcpyHostToDevice);
    static int next_slot = 0;
    e->model_slot = (next_slot++) % 2;   // up to 2 live scenes per process share the constant bank round-robin
    if (err == cudaSuccess) err = mopa::env_slot_claim(e, nullptr, true);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err == cudaSuccess) {   // tuning / profiling hooks of the warp kernel (defaults: no profiling, all stage barriers)
        const char *pf = getenv("MOPA_ENV_PROF"), *sm = getenv("MOPA_ENV_SYNC_MASK");
        err = mopa::env_tune_set(pf ? atoi(pf) : 0, sm ? (int)strtol(sm, nullptr, 0) : 0xFE);
    }
    if (err != cudaSuccess) {
        mopa_set_error(std::string("mopa_env_create: ") + cudaGetErrorString(err) + " (a CUDA device is required; there is no CPU fallback)");
        delete e;
        return MOPA_ERR_CUDA;
    }
    *out = e;
    return MOPA_OK;
}

void mopa_env_destroy(mopa_env *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    mopa::env_slot_release(e);
    if (e->d_model) cudaFree(e->d_model);
    if (e->d_qpos0) cudaFree(e->d_qpos0);
    delete e;
}

int mopa_env_enable_contacts(mopa_env *e, int32_t on) {
    if (!e) return MOPA_ERR_ARG;
    e->h_model.enable_contacts = on ? 1 : 0;
    ENV_TRY(cudaSetDevice(e->device));
    ENV_TRY(cudaMemcpy(e->d_model, &e->h_model, sizeof(mopa::DynDev), cudaMemcpyHostToDevice));
    ENV_TRY(mopa::env_slot_claim(e, nullptr, true));
    ENV_TRY(cudaDeviceSynchronize());
    return MOPA_OK;
}

int mopa_env_debug_prof(mopa_env *e, uint64_t *out32) {
    if (!e || !out32) return MOPA_ERR_ARG;
    ENV_TRY(cudaSetDevice(e->device));
    ENV_TRY(cudaDeviceSynchronize());
    ENV_TRY(mopa::env_prof_read((unsigned long long *)out32));
    return MOPA_OK;
}

int mopa_env_forward(mopa_env *e, const mopa_env_buffers *buf, const int32_t *d_ids, int32_t n, void *stream) {
    if (!e || !buf || n < 0) { mopa_set_error("mopa_env_forward: bad argument"); return MOPA_ERR_ARG; }
    if (n == 0) return MOPA_OK;
    ENV_TRY(cudaSetDevice(e->device));
    ENV_TRY(mopa::launch_env_warp(e, *buf, nullptr, 0, nullptr, nullptr, n, 1, d_ids, (cudaStream_t)stream));
    return MOPA_OK;
}

int mopa_env_reset_pusher(mopa_env *e, const mopa_env_buffers *buf, const uint8_t *d_mask, uint64_t seed, int64_t env_id_offset,
                          int64_t *d_episode, const double *qpos0, int32_t n_envs, void *stream) {
    if (!e || !buf || !d_episode || !qpos0 || n_envs < 0 || e->task.kind != 3) { mopa_set_error("mopa_env_reset_pusher: bad argument"); return MOPA_ERR_ARG; }
    if (n_envs == 0) return MOPA_OK;
    ENV_TRY(cudaSetDevice(e->device));
    if (!e->d_qpos0) {
        ENV_TRY(cudaMalloc(&e->d_qpos0, sizeof(double) * e->h_model.nq));
        ENV_TRY(cudaMemcpy(e->d_qpos0, qpos0, sizeof(double) * e->h_model.nq, cudaMemcpyHostToDevice));
    }
    ENV_TRY(mopa::launch_pusher(e, *buf, nullptr, 0, nullptr, d_mask, n_envs, 3, nullptr, (unsigned long long)seed, (long long)env_id_offset, (long long *)d_episode, (cudaStream_t)stream));
    return MOPA_OK;
}

int mopa_env_step(mopa_env *e, const mopa_env_buffers *buf, const float *d_action, int32_t action_stride,
                  const uint8_t *d_is_planner, const uint8_t *d_mask, int32_t n_envs, void *stream) {
    if (!e || !buf || !d_action || action_stride < (e->task.kind == 3 ? 4 : 7) || n_envs < 0) { mopa_set_error("mopa_env_step: bad argument"); return MOPA_ERR_ARG; }
    if (n_envs == 0) return MOPA_OK;
    ENV_TRY(cudaSetDevice(e->device));
    ENV_TRY(mopa::launch_env_warp(e, *buf, d_action, action_stride, d_is_planner, d_mask, n_envs, 0, nullptr, (cudaStream_t)stream));
    return MOPA_OK;
}

}  // extern "C"
