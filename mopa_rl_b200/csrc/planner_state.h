// Planner handle behind the C ABI (include/mopa_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "../../include/mopa_model_desc.h"
#include "scene.h"

namespace mopa {

constexpr int PLAN_MAXD = 8;  // active joints per planner (Sawyer 7, Pusher 4)
#define MOPA_PLAN_NOT_EXACT_ (-4)
#define MOPA_PLAN_INVALID_GOAL_ (-5)

// The OMPL state space the reference builds over the active joints
// (makeCompoundStateSpace, motion_planners/src/mujoco_ompl_interface.cpp:149-281): one
// weight-1 subspace per joint, R^1 with jnt_range bounds for limited hinge/slide joints,
// SO(2) for unlimited hinges.
struct Space {
    int nq = 0, n_active = 0;
    std::vector<int> active_qadr;
    std::vector<float> lo, hi;
    std::vector<int> is_so2;
    float range = 0.f, resolution = 0.005f;
    uint64_t seed = 0;
};

void build_space(const mopa_model_desc *d, const int32_t *passive, int n_passive, double range, double resolution,
                 uint64_t seed, Space &out);

}  // namespace mopa

struct mopa_planner {
    int device = 0, sm_count = 148;
    mopa::HostScene scene;
    mopa::Space space;
    unsigned char *d_blob = nullptr;
    cudaStream_t stream = nullptr;
    // staging for the *_host entry points (pinned host + device)
    size_t stage_cap = 0;
    float *d_stage_q = nullptr, *h_stage_q = nullptr;
    uint32_t *d_stage_r = nullptr, *h_stage_r = nullptr;
    // double-buffered pipeline of mopa_is_valid_host_f32
    float *pipe_q[2] = {nullptr, nullptr};
    uint32_t *pipe_r[2] = {nullptr, nullptr};
    cudaStream_t pipe_stream[2] = {nullptr, nullptr};
    int pipe_stride = 0;
    // mopa_is_valid_active_host_f32: compact rows (active joints only) as uploaded, and the row the passive joints come from
    float *pipe_a[2] = {nullptr, nullptr};
    float *d_base_row = nullptr;
    int *d_active_adr = nullptr;
    // RRT-Connect work buffers (plan.cu)
    void *plan_buffers = nullptr;
    int max_nodes = 4096;  // node capacity per tree
};

#include <string>
namespace mopa {
void free_plan_buffers(mopa_planner *p);
cudaError_t launch_plan(mopa_planner *p, const float *d_start, const float *d_goal, int row_stride, const unsigned long long *d_keys,
                        int n, int max_iter, float *d_path, int *d_node_ids, int max_path, int *d_path_len, int *d_status,
                        int *d_iters, int *d_nodes, cudaStream_t stream, const int *d_n = nullptr, int cta_warps = 8, float range_override = 0.f,
                        int only_failed = 0, unsigned long long key_xor = 0ULL);
// range_override > 0: extension range of this launch (the "simple" planner of SACAgent, rl/sac_agent.py:98-110, shares the scene
// and state space of the main one); only_failed: skip problems whose d_status is already MOPA_PLAN_OK (retry pass);
// key_xor: mixed into every problem key (independent sample streams for the retry).
int plan_host(mopa_planner *p, const double *start, const double *goal, const uint64_t *keys, int n, int max_iter, double *path,
              int32_t *node_ids, int max_path, int32_t *path_len, int32_t *status, int32_t *iters, std::string &err);
}
