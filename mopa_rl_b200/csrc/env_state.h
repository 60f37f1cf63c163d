// Environment handle behind the C ABI (include/mopa_b200.h) and the launchers of the env-step kernel.
#pragma once
#include <cuda_runtime.h>

#include "../../include/mopa_b200.h"
#include "dyn.cuh"

namespace mopa {
struct DynDev;
cudaError_t env_tune_set(int prof, int sync_mask);
cudaError_t env_prof_read(unsigned long long *out);
}

struct mopa_env {
    int device = 0, sm_count = 148;
    int model_slot = 0;          // slot of this scene in the warp kernel's constant memory
    mopa::DynDev *d_model = nullptr;
    mopa::DynDev h_model;
    mopa_sawyer_task task;
    double *d_qpos0 = nullptr;   // pusher: keyframe the reset noise is added to
};

namespace mopa {
// env.step / sim.forward launch for the handle's scene (refreshes its constant-memory slot when another handle used it)
cudaError_t launch_env_warp(mopa_env *env, const mopa_env_buffers &B, const float *action, int action_stride, const uint8_t *is_planner,
                            const uint8_t *mask, int n, int forward_only, const int32_t *ids, cudaStream_t stream);
cudaError_t env_slot_claim(mopa_env *env, cudaStream_t stream, bool force);
// PusherObstacle-v0: fwd 0 env.step, 1 sim.forward + observation, 3 _reset (rejection sampling on the device)
cudaError_t launch_pusher(mopa_env *env, const mopa_env_buffers &B, const float *action, int action_stride, const uint8_t *is_planner,
                          const uint8_t *mask, int n, int fwd, const int32_t *ids, unsigned long long seed, long long env_id_offset,
                          long long *d_episode, cudaStream_t stream);
void env_slot_release(const mopa_env *env);
}
