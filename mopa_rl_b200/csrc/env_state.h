// Environment handle behind the C ABI (include/mopa_b200.h) and the launchers of the env-step kernel.
#pragma once
#include <cuda_runtime.h>

#include "../../include/mopa_b200.h"
#include "dyn.cuh"

namespace mopa {
struct DynDev;
cudaError_t upload_env_model(int slot, const DynDev &h_model);
cudaError_t env_tune_set(int prof, int sync_mask);
cudaError_t env_prof_read(unsigned long long *out);
cudaError_t launch_env_warp(int model_slot, const DynDev *d_model, int nb, int ngeom, int ngm, const mopa_sawyer_task &T, const mopa_env_buffers &B, const float *action,
                            int action_stride, const uint8_t *is_planner, const uint8_t *mask, int n, int forward_only,
                            const int32_t *ids, cudaStream_t stream);
}

struct mopa_env {
    int device = 0;
    int model_slot = 0;          // slot of this scene in the warp kernel's constant memory
    mopa::DynDev *d_model = nullptr;
    mopa::DynDev h_model;
    mopa_sawyer_task task;
};
