// Batched inverse kinematics: damped least squares on a site pose, one thread per problem.
//
// Replaces qpos_from_site_pose (env/inverse_kinematics.py:18-135; caller MoPARolloutRunner._cart2dispalcement,
// rl/mopa_rollouts.py:683-728): up to max_steps iterations of
//     err = target - site pose;  stop when |err| < tol
//     dq  = (J^T J + lambda I)^-1 J^T err        (nullspace_method with regularization_strength = 3e-2, always on)
//     stop when |err| / |dq| > progress_thresh;  |dq| capped at max_update_norm;  q[arm] += dq
// on a private copy of qpos (the reference drives a second "ik_env" through set_state).  The site Jacobian comes from
// the same kinematic chain the env-step kernel uses (hinge / slide joints on the path from the tree root to the body
// that carries the site).  fp64 throughout, like the reference.
#include <cuda_runtime.h>

#include <string>

#include "../../include/mopa_b200.h"
#include "env_state.h"

void mopa_set_error(const std::string &s);

namespace mopa {

constexpr int IK_MAXCHAIN = 16;   // bodies from the tree root to the site's body
constexpr int IK_MAXJ = 8;        // movable joints

struct IkArgs {
    int n, nq, body, nj, max_steps, use_quat;
    int jdof[IK_MAXJ];            // simulated-dof index of every movable joint
    double site[3], tol, rot_weight, max_update_norm, progress_thresh, reg;
    // rollout mode (MoPARolloutRunner._cart2dispalcement): the target comes from the policy's Cartesian action and the start pose
    int roll;
    double action_range, world_lo[3], world_hi[3], jlo[IK_MAXJ], jhi[IK_MAXJ];
};

// mju_mat2Quat / mju_quat2Vel (dt = 1) as the reference calls them through dm_control's mjlib
__device__ inline void ik_mat2quat(double *q, const double *m) {
    if (m[0] + m[4] + m[8] > 0) {
        q[0] = 0.5 * sqrt(1 + m[0] + m[4] + m[8]);
        q[1] = 0.25 * (m[7] - m[5]) / q[0]; q[2] = 0.25 * (m[2] - m[6]) / q[0]; q[3] = 0.25 * (m[3] - m[1]) / q[0];
    } else if (m[0] > m[4] && m[0] > m[8]) {
        q[1] = 0.5 * sqrt(1 + m[0] - m[4] - m[8]);
        q[0] = 0.25 * (m[7] - m[5]) / q[1]; q[2] = 0.25 * (m[1] + m[3]) / q[1]; q[3] = 0.25 * (m[2] + m[6]) / q[1];
    } else if (m[4] > m[8]) {
        q[2] = 0.5 * sqrt(1 - m[0] + m[4] - m[8]);
        q[0] = 0.25 * (m[2] - m[6]) / q[2]; q[1] = 0.25 * (m[1] + m[3]) / q[2]; q[3] = 0.25 * (m[5] + m[7]) / q[2];
    } else {
        q[3] = 0.5 * sqrt(1 - m[0] - m[4] + m[8]);
        q[0] = 0.25 * (m[3] - m[1]) / q[3]; q[1] = 0.25 * (m[2] + m[6]) / q[3]; q[2] = 0.25 * (m[5] + m[7]) / q[3];
    }
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int k = 0; k < 4; k++) q[k] /= n;
}
__device__ inline void ik_quat2vel(double *res, const double *q) {
    double ax[3] = {q[1], q[2], q[3]};
    const double s = sqrt(d_dot(ax, ax));
    if (s > 0) for (int k = 0; k < 3; k++) ax[k] /= s;
    double speed = 2 * atan2(s, q[0]);
    if (speed > 3.141592653589793) speed -= 2 * 3.141592653589793;
    for (int k = 0; k < 3; k++) res[k] = ax[k] * speed;
}

__global__ void ik_kernel(const DynDev *__restrict__ mg, IkArgs A, const double *__restrict__ qpos_in, const double *__restrict__ target_pos,
                          const double *__restrict__ target_quat, double *__restrict__ qpos_out, double *__restrict__ err_out,
                          int *__restrict__ steps_out, unsigned char *__restrict__ success_out,
                          const float *__restrict__ policy_ac, const unsigned char *__restrict__ need, float *__restrict__ joint_ac) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.n) return;
    if (A.roll && need && !need[t]) return;   // rollout mode: only the environments that start a macro action
    const DynDev &m = *mg;
    double tpos[3], tquat[4];   // the target of this problem
    if (!A.roll) {
        for (int k = 0; k < 3; k++) tpos[k] = target_pos[(size_t)t * 3 + k];
        if (A.use_quat) for (int k = 0; k < 4; k++) tquat[k] = target_quat[(size_t)t * 4 + k];
    }
    double *q = qpos_out + (size_t)t * A.nq;
    for (int k = 0; k < A.nq; k++) q[k] = qpos_in[(size_t)t * A.nq + k];
    // chain of bodies from the tree root to the site's body
    int chain[IK_MAXCHAIN], nc = 0;
    for (int b = A.body; b >= 0 && nc < IK_MAXCHAIN; b = m.b_parent[b]) chain[nc++] = b;
    const int ne = A.use_quat ? 6 : 3;
    double err_norm = 0;
    int steps = 0;
    bool success = false;
    for (steps = 0; steps < A.max_steps; steps++) {
        // ---- kinematics along the chain (same composition as the env-step kernel: parent frame o body frame o joint)
        double pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0}, R[9];
        double jax[IK_MAXJ][3], janc[IK_MAXJ][3];   // world axis / anchor of the movable joints met on the chain
        int jhit[IK_MAXJ], jtype[IK_MAXJ];
        for (int j = 0; j < A.nj; j++) jhit[j] = 0;
        for (int c = nc - 1; c >= 0; c--) {
            const int i = chain[c], jt = m.b_jtype[i];
            double Pp[3], Pq[4], PM[9], tv[3];
            if (c == nc - 1) {
                for (int k = 0; k < 3; k++) Pp[k] = m.b_rootpos[i][k];
                for (int k = 0; k < 4; k++) Pq[k] = m.b_rootquat[i][k];
            } else {
                for (int k = 0; k < 3; k++) Pp[k] = pos[k];
                for (int k = 0; k < 4; k++) Pq[k] = quat[k];
            }
            if (jt == 0) {   // free joint: absolute pose
                const int a = m.b_qadr[i];
                for (int k = 0; k < 3; k++) pos[k] = q[a + k];
                const double nn = sqrt(q[a + 3] * q[a + 3] + q[a + 4] * q[a + 4] + q[a + 5] * q[a + 5] + q[a + 6] * q[a + 6]);
                for (int k = 0; k < 4; k++) quat[k] = q[a + 3 + k] / nn;
                continue;
            }
            d_q2m(PM, Pq);
            d_mv(tv, PM, m.b_pos[i]);
            for (int k = 0; k < 3; k++) pos[k] = Pp[k] + tv[k];
            double bq[4];
            d_qmul(bq, Pq, m.b_quat[i]);
            for (int k = 0; k < 4; k++) quat[k] = bq[k];
            if (jt == 3 || jt == 2) {
                d_q2m(R, quat);
                double ax[3], anchor[3];
                d_mv(tv, R, m.b_jpos[i]);
                for (int k = 0; k < 3; k++) anchor[k] = pos[k] + tv[k];
                if (jt == 3) {
                    const double ang = q[m.b_qadr[i]] - m.b_qpos0[i], sn = sin(0.5 * ang), cs = cos(0.5 * ang);
                    const double ql[4] = {cs, sn * m.b_jaxis[i][0], sn * m.b_jaxis[i][1], sn * m.b_jaxis[i][2]};
                    double qn[4];
                    d_qmul(qn, quat, ql);
                    for (int k = 0; k < 4; k++) quat[k] = qn[k];
                    d_q2m(R, quat);
                    d_mv(tv, R, m.b_jpos[i]);
                    for (int k = 0; k < 3; k++) pos[k] = anchor[k] - tv[k];
                    d_mv(ax, R, m.b_jaxis[i]);
                } else {
                    d_mv(ax, R, m.b_jaxis[i]);
                    const double dq = q[m.b_qadr[i]] - m.b_qpos0[i];
                    for (int k = 0; k < 3; k++) pos[k] += ax[k] * dq;
                }
                for (int j = 0; j < A.nj; j++)
                    if (A.jdof[j] == m.b_dadr[i]) {
                        jhit[j] = 1; jtype[j] = jt;
                        for (int k = 0; k < 3; k++) { jax[j][k] = ax[k]; janc[j][k] = anchor[k]; }
                    }
            }
        }
        d_q2m(R, quat);
        double sp[3], tv[3];
        d_mv(tv, R, A.site);
        for (int k = 0; k < 3; k++) sp[k] = pos[k] + tv[k];
        if (A.roll && steps == 0) {
            // rl/mopa_rollouts.py:91-98: target_cart = clip(site_xpos + action_range * ac["default"], min_world_size, max_world_size)
            const float *ac = policy_ac + (size_t)t * 8;
            for (int k = 0; k < 3; k++) {
                const double c = sp[k] + A.action_range * (double)ac[k];
                tpos[k] = c < A.world_lo[k] ? A.world_lo[k] : (c > A.world_hi[k] ? A.world_hi[k] : c);
            }
            // :692-697: util.env.mat2quat(site_xmat) - the unit quaternion of the float32 copy of the matrix with w >= 0, returned as
            // (x, y, z, w) - then the index list [3, 0, 1, 1] (sic) = (w, x, y, y), times the normalised action quaternion
            double Rf[9], sq[4];
            for (int k = 0; k < 9; k++) Rf[k] = (double)(float)R[k];
            ik_mat2quat(sq, Rf);
            if (sq[0] < 0) for (int k = 0; k < 4; k++) sq[k] = -sq[k];
            const double t0[4] = {sq[0], sq[1], sq[2], sq[2]};
            const float a0 = ac[3], a1 = ac[4], a2 = ac[5], a3 = ac[6];
            const float an = sqrtf(a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3);
            const double aq[4] = {(double)(a0 / an), (double)(a1 / an), (double)(a2 / an), (double)(a3 / an)};
            d_qmul(tquat, t0, aq);
        }
        // ---- error
        double err[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 3; k++) err[k] = tpos[k] - sp[k];
        err_norm = sqrt(err[0] * err[0] + err[1] * err[1] + err[2] * err[2]);
        if (A.use_quat) {
            double sq[4], nq4[4], eq[4];
            ik_mat2quat(sq, R);
            nq4[0] = sq[0]; nq4[1] = -sq[1]; nq4[2] = -sq[2]; nq4[3] = -sq[3];
            d_qmul(eq, tquat, nq4);
            ik_quat2vel(err + 3, eq);
            err_norm += sqrt(err[3] * err[3] + err[4] * err[4] + err[5] * err[5]) * A.rot_weight;
        }
        if (err_norm < A.tol) { success = true; break; }
        // ---- site Jacobian columns of the movable joints (get_site_jacp / get_site_jacr)
        double J[6][IK_MAXJ];
        for (int j = 0; j < A.nj; j++) {
            double cp[3] = {0, 0, 0}, cr[3] = {0, 0, 0};
            if (jhit[j]) {
                if (jtype[j] == 3) {
                    const double d[3] = {sp[0] - janc[j][0], sp[1] - janc[j][1], sp[2] - janc[j][2]};
                    d_cross(cp, jax[j], d);
                    for (int k = 0; k < 3; k++) cr[k] = jax[j][k];
                } else
                    for (int k = 0; k < 3; k++) cp[k] = jax[j][k];
            }
            for (int k = 0; k < 3; k++) { J[k][j] = cp[k]; J[3 + k][j] = cr[k]; }
        }
        // ---- dq = (J^T J + reg I)^-1 J^T err  (symmetric positive definite: Cholesky)
        double H[IK_MAXJ][IK_MAXJ], g[IK_MAXJ];
        for (int a = 0; a < A.nj; a++) {
            double s = 0;
            for (int k = 0; k < ne; k++) s += J[k][a] * err[k];
            g[a] = s;
            for (int b = 0; b <= a; b++) {
                double h = 0;
                for (int k = 0; k < ne; k++) h += J[k][a] * J[k][b];
                H[a][b] = h + (a == b ? A.reg : 0.0);
            }
        }
        for (int a = 0; a < A.nj; a++)
            for (int b = 0; b <= a; b++) {
                double s = H[a][b];
                for (int k = 0; k < b; k++) s -= H[a][k] * H[b][k];
                H[a][b] = (a == b) ? sqrt(s) : s / H[b][b];
            }
        for (int a = 0; a < A.nj; a++) { double s = g[a]; for (int k = 0; k < a; k++) s -= H[a][k] * g[k]; g[a] = s / H[a][a]; }
        for (int a = A.nj - 1; a >= 0; a--) { double s = g[a]; for (int k = a + 1; k < A.nj; k++) s -= H[k][a] * g[k]; g[a] = s / H[a][a]; }
        double un = 0;
        for (int a = 0; a < A.nj; a++) un += g[a] * g[a];
        un = sqrt(un);
        if (err_norm / un > A.progress_thresh) break;
        const double sc = un > A.max_update_norm ? A.max_update_norm / un : 1.0;
        for (int j = 0; j < A.nj; j++) q[m.d_qadr[A.jdof[j]]] += g[j] * sc;
    }
    if (A.roll) {
        // :710-727: target_qpos[ref] = result.qpos[ref], clipped to the joint ranges; displacement = target - current (+ the gripper entry)
        for (int j = 0; j < A.nj; j++) {
            const int a = m.d_qadr[A.jdof[j]];
            double v = q[a];
            v = v < A.jlo[j] ? A.jlo[j] : (v > A.jhi[j] ? A.jhi[j] : v);
            joint_ac[(size_t)t * 8 + j] = (float)(v - qpos_in[(size_t)t * A.nq + a]);
        }
        joint_ac[(size_t)t * 8 + 7] = policy_ac[(size_t)t * 8 + 7];
        return;
    }
    if (steps == A.max_steps && steps > 0) steps = A.max_steps - 1;   // Python's `for steps in range(max_steps)` leaves the last index
    err_out[t] = err_norm;
    steps_out[t] = steps;
    success_out[t] = success ? 1 : 0;
}

// rollout front end of mopa_rollout_step_ik: Cartesian policy actions -> joint displacement rows for the environments flagged in `need`
cudaError_t launch_ik_rollout(mopa_env *e, const double *d_qpos, const float *d_policy_ac, const unsigned char *d_need, int body,
                              const double *site_local, const int *joint_dofs, int n_joints, int n, int max_steps, double tol,
                              double action_range, const double *world_lo, const double *world_hi, const double *jlo, const double *jhi,
                              double *d_qpos_scratch, float *d_joint_ac, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if (n_joints < 1 || n_joints > IK_MAXJ - 1) return cudaErrorInvalidValue;
    IkArgs A;
    A.n = n; A.nq = e->h_model.nq; A.body = body; A.nj = n_joints; A.max_steps = max_steps; A.use_quat = 1; A.roll = 1;
    for (int j = 0; j < n_joints; j++) { A.jdof[j] = joint_dofs[j]; A.jlo[j] = jlo[j]; A.jhi[j] = jhi[j]; }
    for (int k = 0; k < 3; k++) { A.site[k] = site_local[k]; A.world_lo[k] = world_lo[k]; A.world_hi[k] = world_hi[k]; }
    A.tol = tol; A.rot_weight = 1.0; A.max_update_norm = 2.0; A.progress_thresh = 20.0; A.reg = 3e-2; A.action_range = action_range;
    ik_kernel<<<(n + 63) / 64, 64, 0, stream>>>(e->d_model, A, d_qpos, nullptr, nullptr, d_qpos_scratch, nullptr, nullptr, nullptr, d_policy_ac, d_need,
                                                d_joint_ac);
    return cudaGetLastError();
}

}  // namespace mopa

extern "C" int mopa_ik_batch(mopa_env *e, const double *d_qpos, const double *d_target_pos, const double *d_target_quat, int32_t body,
                             const double *site_local, const int32_t *joint_dofs, int32_t n_joints, int32_t n, int32_t max_steps, double tol,
                             double *d_qpos_out, double *d_err, int32_t *d_steps, uint8_t *d_success, void *stream) {
    if (!e || !d_qpos || !d_target_pos || !site_local || !joint_dofs || !d_qpos_out || !d_err || !d_steps || !d_success || n < 0 ||
        n_joints < 1 || n_joints > mopa::IK_MAXJ || body < 0 || body >= e->h_model.nb) {
        mopa_set_error("mopa_ik_batch: bad argument");
        return MOPA_ERR_ARG;
    }
    if (n == 0) return MOPA_OK;
    mopa::IkArgs A;
    A.roll = 0;
    A.n = n; A.nq = e->h_model.nq; A.body = body; A.nj = n_joints; A.max_steps = max_steps; A.use_quat = d_target_quat ? 1 : 0;
    for (int j = 0; j < n_joints; j++) A.jdof[j] = joint_dofs[j];
    for (int k = 0; k < 3; k++) A.site[k] = site_local[k];
    A.tol = tol; A.rot_weight = 1.0; A.max_update_norm = 2.0; A.progress_thresh = 20.0; A.reg = 3e-2;   // inverse_kinematics.py:22-33
    cudaError_t err = cudaSetDevice(e->device);
    if (err == cudaSuccess) {
        mopa::ik_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(e->d_model, A, d_qpos, d_target_pos, d_target_quat, d_qpos_out, d_err, d_steps,
                                                                        d_success, nullptr, nullptr, nullptr);
        err = cudaGetLastError();
    }
    if (err != cudaSuccess) { mopa_set_error(std::string("mopa_ik_batch: ") + cudaGetErrorString(err)); return MOPA_ERR_CUDA; }
    return MOPA_OK;
}
