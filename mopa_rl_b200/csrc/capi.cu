// extern "C" surface of libmopa_b200.so (declared in include/mopa_b200.h).
#include <cuda_runtime.h>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mopa_b200.h"
#include "planner_state.h"

namespace mopa {
void build_scene(const mopa_model_desc *d, const int32_t *ignored, int nignored, double threshold, HostScene &out);
cudaError_t launch_is_valid(const unsigned char *d_blob, const SceneHeader &H, const float *d_qpos, int row_stride, int n,
                            uint32_t *d_out, int exact, int sm_count, cudaStream_t stream, const int *d_n = nullptr, int d_n_mult = 1);
}  // namespace mopa

static thread_local std::string g_err;
void mopa_set_error(const std::string &s) { g_err = s; }
#define CUDA_TRY(x)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) {                                                                 \
            g_err = std::string(#x) + ": " + cudaGetErrorString(e_);                             \
            return MOPA_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

extern "C" {

const char *mopa_last_error(void) { return g_err.c_str(); }

int mopa_abi_sizes(int32_t *out5) {
    if (!out5) { g_err = "mopa_abi_sizes: bad argument"; return MOPA_ERR_ARG; }
    out5[0] = (int32_t)sizeof(mopa_model_desc); out5[1] = (int32_t)sizeof(mopa_dyn_desc); out5[2] = (int32_t)sizeof(mopa_sawyer_task);
    out5[3] = (int32_t)sizeof(mopa_env_buffers); out5[4] = (int32_t)sizeof(mopa_rollout_config);
    return MOPA_OK;
}

int mopa_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        return MOPA_ERR_CUDA;
    }
    return n;
}

int mopa_planner_create(const mopa_model_desc *model, const int32_t *passive_qpos_idx, int32_t n_passive,
                        const int32_t *ignored_pairs, int32_t n_ignored, double contact_threshold, double range,
                        double resolution, uint64_t seed, int32_t device, mopa_planner **out) {
    if (!model || !out || n_passive < 0 || n_ignored < 0) { g_err = "mopa_planner_create: bad argument"; return MOPA_ERR_ARG; }
    *out = nullptr;
    mopa_planner *p = new mopa_planner();
    try {
        mopa::build_scene(model, ignored_pairs, n_ignored, contact_threshold, p->scene);
        mopa::build_space(model, passive_qpos_idx, n_passive, range, resolution, seed, p->space);
    } catch (const std::exception &e) {
        g_err = std::string("mopa_planner_create: ") + e.what();
        delete p;
        return MOPA_ERR_MODEL;
    }
    p->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_blob, p->scene.blob.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->d_blob, p->scene.blob.data(), p->scene.blob.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        g_err = std::string("mopa_planner_create: ") + cudaGetErrorString(e) +
                " (libmopa_b200 needs a CUDA device; there is no CPU fallback)";
        mopa_planner_destroy(p);
        return MOPA_ERR_CUDA;
    }
    *out = p;
    return MOPA_OK;
}

void mopa_planner_destroy(mopa_planner *p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->d_blob) cudaFree(p->d_blob);
    if (p->d_stage_q) cudaFree(p->d_stage_q);
    if (p->d_stage_r) cudaFree(p->d_stage_r);
    if (p->h_stage_q) cudaFreeHost(p->h_stage_q);
    if (p->h_stage_r) cudaFreeHost(p->h_stage_r);
    mopa::free_plan_buffers(p);
    for (int k = 0; k < 2; k++) {
        if (p->pipe_q[k]) cudaFree(p->pipe_q[k]);
        if (p->pipe_r[k]) cudaFree(p->pipe_r[k]);
        if (p->pipe_a[k]) cudaFree(p->pipe_a[k]);
        if (p->pipe_stream[k]) cudaStreamDestroy(p->pipe_stream[k]);
    }
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

int mopa_planner_info(const mopa_planner *p, int32_t *nq, int32_t *n_pairs, int32_t *n_active) {
    if (!p) { g_err = "null planner"; return MOPA_ERR_ARG; }
    if (nq) *nq = p->scene.hdr.nq;
    if (n_pairs) *n_pairs = (int32_t)p->scene.canon_g1.size();
    if (n_active) *n_active = p->space.n_active;
    return MOPA_OK;
}

int mopa_scene_pair_table(const mopa_model_desc *model, const int32_t *ignored_pairs, int32_t n_ignored, double contact_threshold,
                          int32_t *stats, uint8_t *kept, int32_t n_pairs) {
    if (!model || !stats) { g_err = "bad argument"; return MOPA_ERR_ARG; }
    try {
        mopa::HostScene sc;
        mopa::build_scene(model, ignored_pairs, n_ignored, contact_threshold, sc);
        const mopa::SceneHeader &H = sc.hdr;
        const int32_t st[8] = {H.n_pair, H.n_real, H.n_group, H.n_pruned, (int32_t)sc.canon_g1.size(), H.blob_bytes, H.frame_floats, 0};
        for (int k = 0; k < 8; k++) stats[k] = st[k];
        if (kept) {
            if (n_pairs != (int32_t)sc.canon_g1.size()) { g_err = "n_pairs does not match the canonical pair count"; return MOPA_ERR_ARG; }
            for (int i = 0; i < n_pairs; i++) kept[i] = 0;
            const mopa::PairRec *pairs = reinterpret_cast<const mopa::PairRec *>(sc.blob.data() + H.off_pair);
            const uint16_t *real = reinterpret_cast<const uint16_t *>(sc.blob.data() + H.off_real);
            for (int i = 0; i < H.n_real; i++) kept[pairs[real[i]].canon] = 1;
        }
    } catch (const std::exception &e) { g_err = e.what(); return MOPA_ERR_ARG; }
    return MOPA_OK;
}

int mopa_planner_pairs(const mopa_planner *p, int32_t *geom1, int32_t *geom2) {
    if (!p || !geom1 || !geom2) { g_err = "bad argument"; return MOPA_ERR_ARG; }
    for (size_t i = 0; i < p->scene.canon_g1.size(); i++) { geom1[i] = p->scene.canon_g1[i]; geom2[i] = p->scene.canon_g2[i]; }
    return MOPA_OK;
}

int mopa_is_valid_batch(mopa_planner *p, const float *d_qpos, int32_t row_stride, int32_t n, uint32_t *d_result,
                        int32_t flags, void *stream) {
    if (!p || n < 0 || (n > 0 && (!d_qpos || !d_result)) || row_stride < p->scene.hdr.nq) {
        g_err = "mopa_is_valid_batch: bad argument";
        return MOPA_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(mopa::launch_is_valid(p->d_blob, p->scene.hdr, d_qpos, row_stride, n, d_result, flags & MOPA_VALID_FIRST_PAIR,
                                   p->sm_count, (cudaStream_t)stream));
    return MOPA_OK;
}

static int ensure_stage(mopa_planner *p, size_t n) {
    if (n <= p->stage_cap) return MOPA_OK;
    size_t cap = p->stage_cap ? p->stage_cap : 256;
    while (cap < n) cap *= 2;
    if (p->d_stage_q) cudaFree(p->d_stage_q);
    if (p->d_stage_r) cudaFree(p->d_stage_r);
    if (p->h_stage_q) cudaFreeHost(p->h_stage_q);
    if (p->h_stage_r) cudaFreeHost(p->h_stage_r);
    p->d_stage_q = nullptr; p->d_stage_r = nullptr; p->h_stage_q = nullptr; p->h_stage_r = nullptr;
    p->stage_cap = 0;
    const size_t row = (size_t)p->scene.hdr.nq4 * 4;
    CUDA_TRY(cudaMalloc(&p->d_stage_q, cap * row * sizeof(float)));
    CUDA_TRY(cudaMalloc(&p->d_stage_r, cap * sizeof(uint32_t)));
    CUDA_TRY(cudaMallocHost(&p->h_stage_q, cap * row * sizeof(float)));
    CUDA_TRY(cudaMallocHost(&p->h_stage_r, cap * sizeof(uint32_t)));
    p->stage_cap = cap;
    return MOPA_OK;
}

int mopa_is_valid_host(mopa_planner *p, const double *qpos, int32_t n, uint8_t *valid, uint32_t *words, int32_t flags) {
    if (!p || n < 0 || (n > 0 && (!qpos || !valid))) { g_err = "mopa_is_valid_host: bad argument"; return MOPA_ERR_ARG; }
    if (n == 0) return MOPA_OK;
    CUDA_TRY(cudaSetDevice(p->device));
    int rc = ensure_stage(p, (size_t)n);
    if (rc) return rc;
    const int nq = p->scene.hdr.nq, row = p->scene.hdr.nq4 * 4;
    for (int i = 0; i < n; i++) {
        float *dst = p->h_stage_q + (size_t)i * row;
        const double *src = qpos + (size_t)i * nq;
        for (int k = 0; k < nq; k++) dst[k] = (float)src[k];
        for (int k = nq; k < row; k++) dst[k] = 0.0f;
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_stage_q, p->h_stage_q, (size_t)n * row * sizeof(float), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(mopa::launch_is_valid(p->d_blob, p->scene.hdr, p->d_stage_q, row, n, p->d_stage_r, flags & MOPA_VALID_FIRST_PAIR,
                                   p->sm_count, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->h_stage_r, p->d_stage_r, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    for (int i = 0; i < n; i++) {
        valid[i] = (uint8_t)(p->h_stage_r[i] & 1u);
        if (words) words[i] = p->h_stage_r[i];
    }
    return MOPA_OK;
}

int mopa_is_valid_host_f32(mopa_planner *p, const float *qpos, int32_t row_stride, int32_t n, uint32_t *words, int32_t flags) {
    if (!p || n < 0 || row_stride < p->scene.hdr.nq || (n > 0 && (!qpos || !words))) { g_err = "mopa_is_valid_host_f32: bad argument"; return MOPA_ERR_ARG; }
    if (n == 0) return MOPA_OK;
    CUDA_TRY(cudaSetDevice(p->device));
    const int chunk = 1 << 19;
    if (!p->pipe_q[0] || p->pipe_stride != row_stride) {
        for (int k = 0; k < 2; k++) {
            if (p->pipe_q[k]) cudaFree(p->pipe_q[k]);
            if (p->pipe_r[k]) cudaFree(p->pipe_r[k]);
            p->pipe_q[k] = nullptr; p->pipe_r[k] = nullptr;
            CUDA_TRY(cudaMalloc(&p->pipe_q[k], (size_t)chunk * row_stride * sizeof(float)));
            CUDA_TRY(cudaMalloc(&p->pipe_r[k], (size_t)chunk * sizeof(uint32_t)));
            if (!p->pipe_stream[k]) CUDA_TRY(cudaStreamCreateWithFlags(&p->pipe_stream[k], cudaStreamNonBlocking));
        }
        p->pipe_stride = row_stride;
    }
    int k = 0;
    for (int off = 0; off < n; off += chunk, k ^= 1) {
        const int m = (n - off) < chunk ? (n - off) : chunk;
        cudaStream_t st = p->pipe_stream[k];
        CUDA_TRY(cudaMemcpyAsync(p->pipe_q[k], qpos + (size_t)off * row_stride, (size_t)m * row_stride * sizeof(float), cudaMemcpyHostToDevice, st));
        CUDA_TRY(mopa::launch_is_valid(p->d_blob, p->scene.hdr, p->pipe_q[k], row_stride, m, p->pipe_r[k], flags & MOPA_VALID_FIRST_PAIR, p->sm_count, st));
        CUDA_TRY(cudaMemcpyAsync(words + off, p->pipe_r[k], (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(p->pipe_stream[0]));
    CUDA_TRY(cudaStreamSynchronize(p->pipe_stream[1]));
    return MOPA_OK;
}

// full qpos rows from rows that hold the active joints only: row[adr[k]] = active[k], every other entry from the base row
__global__ void expand_active_rows_kernel(const float *__restrict__ active, int na, const int *__restrict__ adr, const float *__restrict__ base,
                                          int nq, int row_stride, int n, float *__restrict__ rows) {
    const long long total = (long long)n * row_stride;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / row_stride), c = (int)(i - (long long)r * row_stride);
        float v = c < nq ? base[c] : 0.0f;
        for (int k = 0; k < na; k++) if (adr[k] == c) v = active[(size_t)r * na + k];
        rows[i] = v;
    }
}

int mopa_is_valid_active_host_f32(mopa_planner *p, const float *active, int32_t n, const float *base_qpos, uint32_t *words, int32_t flags) {
    if (!p || n < 0 || !base_qpos || (n > 0 && (!active || !words))) { g_err = "mopa_is_valid_active_host_f32: bad argument"; return MOPA_ERR_ARG; }
    if (n == 0) return MOPA_OK;
    CUDA_TRY(cudaSetDevice(p->device));
    const int chunk = 1 << 19, nq = p->scene.hdr.nq, row = 4 * p->scene.hdr.nq4, na = p->space.n_active;
    if (!p->pipe_q[0] || p->pipe_stride != row) {
        for (int k = 0; k < 2; k++) {
            if (p->pipe_q[k]) cudaFree(p->pipe_q[k]);
            if (p->pipe_r[k]) cudaFree(p->pipe_r[k]);
            p->pipe_q[k] = nullptr; p->pipe_r[k] = nullptr;
            CUDA_TRY(cudaMalloc(&p->pipe_q[k], (size_t)chunk * row * sizeof(float)));
            CUDA_TRY(cudaMalloc(&p->pipe_r[k], (size_t)chunk * sizeof(uint32_t)));
            if (!p->pipe_stream[k]) CUDA_TRY(cudaStreamCreateWithFlags(&p->pipe_stream[k], cudaStreamNonBlocking));
        }
        p->pipe_stride = row;
    }
    if (!p->pipe_a[0]) {
        for (int k = 0; k < 2; k++) CUDA_TRY(cudaMalloc(&p->pipe_a[k], (size_t)chunk * na * sizeof(float)));
        CUDA_TRY(cudaMalloc(&p->d_base_row, (size_t)nq * sizeof(float)));
        CUDA_TRY(cudaMalloc(&p->d_active_adr, (size_t)na * sizeof(int)));
        CUDA_TRY(cudaMemcpy(p->d_active_adr, p->space.active_qadr.data(), (size_t)na * sizeof(int), cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMemcpy(p->d_base_row, base_qpos, (size_t)nq * sizeof(float), cudaMemcpyHostToDevice));   // synchronous: ordered before both streams' work
    int k = 0;
    for (int off = 0; off < n; off += chunk, k ^= 1) {
        const int m = (n - off) < chunk ? (n - off) : chunk;
        cudaStream_t st = p->pipe_stream[k];
        CUDA_TRY(cudaMemcpyAsync(p->pipe_a[k], active + (size_t)off * na, (size_t)m * na * sizeof(float), cudaMemcpyHostToDevice, st));
        expand_active_rows_kernel<<<p->sm_count * 8, 256, 0, st>>>(p->pipe_a[k], na, p->d_active_adr, p->d_base_row, nq, row, m, p->pipe_q[k]);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(mopa::launch_is_valid(p->d_blob, p->scene.hdr, p->pipe_q[k], row, m, p->pipe_r[k], flags & MOPA_VALID_FIRST_PAIR, p->sm_count, st));
        CUDA_TRY(cudaMemcpyAsync(words + off, p->pipe_r[k], (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(p->pipe_stream[0]));
    CUDA_TRY(cudaStreamSynchronize(p->pipe_stream[1]));
    return MOPA_OK;
}

int mopa_planner_set_max_nodes(mopa_planner *p, int32_t max_nodes) {
    if (!p || max_nodes < 2) { g_err = "mopa_planner_set_max_nodes: bad argument"; return MOPA_ERR_ARG; }
    p->max_nodes = max_nodes;
    return MOPA_OK;
}

int mopa_plan_batch(mopa_planner *p, const float *d_start, const float *d_goal, int32_t row_stride, const uint64_t *d_keys,
                    int32_t n, int32_t max_iter, float *d_path, int32_t *d_node_ids, int32_t max_path, int32_t *d_path_len,
                    int32_t *d_status, int32_t *d_iters, int32_t *d_nodes, void *stream) {
    if (!p || n < 0 || max_path < 2 || row_stride < p->scene.hdr.nq ||
        (n > 0 && (!d_start || !d_goal || !d_keys || !d_path || !d_node_ids || !d_path_len || !d_status))) {
        g_err = "mopa_plan_batch: bad argument";
        return MOPA_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(mopa::launch_plan(p, d_start, d_goal, row_stride, (const unsigned long long *)d_keys, n, max_iter, d_path, d_node_ids,
                               max_path, d_path_len, d_status, d_iters, d_nodes, (cudaStream_t)stream));
    return MOPA_OK;
}

int mopa_plan_host(mopa_planner *p, const double *start, const double *goal, const uint64_t *keys, int32_t n, int32_t max_iter,
                   double *path, int32_t *node_ids, int32_t max_path, int32_t *path_len, int32_t *status, int32_t *iters) {
    if (!p || n < 0 || max_path < 2 || (n > 0 && (!start || !goal || !keys || !path || !path_len || !status))) {
        g_err = "mopa_plan_host: bad argument";
        return MOPA_ERR_ARG;
    }
    if (n == 0) return MOPA_OK;
    CUDA_TRY(cudaSetDevice(p->device));
    std::string err;
    int rc = mopa::plan_host(p, start, goal, keys, n, max_iter, path, node_ids, max_path, path_len, status, iters, err);
    if (rc) { g_err = "mopa_plan_host: " + err; return MOPA_ERR_CUDA; }
    return MOPA_OK;
}

}  // extern "C"
