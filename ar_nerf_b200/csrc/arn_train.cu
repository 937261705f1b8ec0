// libarnerf.so -- fused training step: the whole forward + backward of one batch of rays issued by ONE host call, with
// the marched sample count kept on the device (no host synchronisation, no Python between kernels).
//   ray/box + near clamp -> march count (+ t record) -> scan -> emit -> hash grid + MLPs (tcgen05) -> composite ->
//   NeRFLoss (rgb "raw" relative L2 + opacity entropy + optional depth term, losses.py:63-82) and its gradient ->
//   composite backward -> MLP + hash-grid backward (gradients accumulated into the caller's buffers).
// The optimizer (arn_adam_step) is a separate call so that the caller can all-reduce the gradients in between.
#include "arn_common.cuh"
#include "arn_field.cuh"

namespace arn {

// losses.py:63-82 for loss_set == 'raw' plus rendering.py:287-296 background blend, forward value and gradient.
// One thread per ray; per-ray results are indexed by ray (rays_a is canonical in the fused step).
__global__ void __launch_bounds__(256) nerf_loss_kernel(const float* __restrict__ rgb, const float* __restrict__ opacity, const float* __restrict__ depth,
                                                        const float* __restrict__ target, int64_t n_rays, float bg_r, float bg_g, float bg_b,
                                                        float lambda_opacity, float lambda_depth, float grid_scale, float grad_scale,
                                                        float* __restrict__ rgb_out, float* __restrict__ dL_drgb, float* __restrict__ dL_dopacity,
                                                        float* __restrict__ dL_ddepth, float* __restrict__ loss_out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.0f;
    if (r < n_rays) {
        const float inv_r = 1.0f / (float)n_rays, inv_3r = inv_r / 3.0f;
        const float o = opacity[r];
        const float bg[3] = {bg_r, bg_g, bg_b};
        float g_op = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float est = rgb[3 * r + c] + bg[c] * (1.0f - o);  // rendering.py:295
            if (rgb_out) rgb_out[3 * r + c] = est;
            const float den = est + 1e-3f;                          // est.detach() + 1e-3
            const float e = (est - target[3 * r + c]) / den;
            loss += e * e * inv_3r;
            const float g = 2.0f * e / den * inv_3r * grad_scale;
            dL_drgb[3 * r + c] = g;
            g_op -= bg[c] * g;
        }
        const float oe = o + 1e-10f;
        loss += lambda_opacity * (-oe * logf(oe)) * inv_r;
        g_op += lambda_opacity * (-logf(oe) - 1.0f) * inv_r * grad_scale;
        dL_dopacity[r] = g_op;
        float g_d = 0.0f;
        if (lambda_depth != 0.0f) {
            const float v = depth[r] / grid_scale + 1e-10f;
            loss += -lambda_depth * logf(fminf(v, 1.0f)) * inv_r;
            if (v < 1.0f) g_d = -lambda_depth / v / grid_scale * inv_r * grad_scale;
        }
        dL_ddepth[r] = g_d;
    }
    // block reduction of the scalar loss
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(kFull, loss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
        if (threadIdx.x == 0) atomicAdd(loss_out, v);
    }
}

}  // namespace arn

using namespace arn;

extern "C" {
int arn_ray_aabb_near(const float*, const float*, int64_t, const float*, const float*, float, float*, arn_stream_t);
int arn_march_train_count_ex(const float*, const float*, const float*, int64_t, const uint8_t*, int, int, float, float, const float*, int, int64_t*,
                             int32_t*, float*, int32_t*, arn_stream_t);
int arn_march_train_emit_dyn(const float*, const float*, int64_t, int, int, float, float, int, const int64_t*, const float*, const int32_t*, float*,
                             float*, float*, float*, int64_t, arn_stream_t);
int arn_composite_train_fw_loss_ex(const float*, const float*, const float*, const float*, const int64_t*, int64_t, int64_t, float, int64_t*, float*, float*,
                                   float*, float*, const float*, const float*, float, float, float, float, float*, float*, float*, float*, float*, int, float*, float*,
                                   arn_stream_t);
int arn_field_fw_tc_dyn(const float*, const float*, int64_t, const int32_t*, const float*, const float*, arn_levels_t, const void*, const void*, int,
                        arn_field_ws_t, float*, float*, arn_stream_t);
int arn_field_bw_tc_dyn(const float*, int64_t, const int32_t*, const float*, const float*, arn_levels_t, const void*, const void*, int, arn_field_ws_t,
                        const float*, const float*, const float*, const float*, float, float*, float*, float*, float*, arn_stream_t);
}

extern "C" ARN_API int arn_nerf_loss(const float* rgb, const float* opacity, const float* depth, const float* target, int64_t n_rays,
                                     const float* bg_host, float lambda_opacity, float lambda_depth, float grid_scale, float grad_scale,
                                     float* rgb_out, float* dL_drgb, float* dL_dopacity, float* dL_ddepth, float* loss_out, arn_stream_t stream) {
    ARN_REQUIRE(n_rays >= 0, "bad size");
    ARN_REQUIRE(loss_out, "null loss_out");
    cudaStream_t st = (cudaStream_t)stream;
    ARN_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    if (n_rays == 0) return ARN_OK;
    ARN_REQUIRE(rgb && opacity && depth && target && bg_host && dL_drgb && dL_dopacity && dL_ddepth, "null pointer");
    ARN_LAUNCH("nerf_loss_kernel", st, nerf_loss_kernel<<<ceil_div(n_rays, 256), 256, 0, st>>>(rgb, opacity, depth, target, n_rays, bg_host[0], bg_host[1],
               bg_host[2], lambda_opacity, lambda_depth, grid_scale, grad_scale, rgb_out, dL_drgb, dL_dopacity, dL_ddepth, loss_out));
    return check_launch("nerf_loss");
}

// The geometry half of the step: depends on the rays and the occupancy bitfield only (not on the weights), so a caller
// may run it for the NEXT batch on a second stream while the current batch is in its field / optimizer kernels.
extern "C" ARN_API int arn_train_march(const arn_train_t* c, arn_stream_t stream) {
    ARN_REQUIRE(c, "null config");
    ARN_REQUIRE(c->n_rays > 0 && c->capacity > 0, "bad sizes");
    const int64_t R = c->n_rays;
    // the step's loss accumulator is zeroed here, with the geometry (on the side stream when the march is prefetched): one
    // node less on the serial chain of arn_train_fwbw_marched.  Every march set has its own accumulator (arn_train_t).
    if (c->loss_out) ARN_CUDA(cudaMemsetAsync(c->loss_out, 0, sizeof(float), (cudaStream_t)stream));
    if (int e = arn_ray_aabb_near(c->rays_o, c->rays_d, R, c->center_host, c->half_size_host, c->near, c->hits_t, stream)) return e;
    if (int e = arn_march_train_count_ex(c->rays_o, c->rays_d, c->hits_t, R, c->density_bitfield, c->cascades, c->grid_size, c->scale,
                                         c->exp_step_factor, c->noise, c->max_samples, c->rays_a, c->counter, c->t_scratch, c->count_scratch, stream)) return e;
    return arn_march_train_emit_dyn(c->rays_o, c->rays_d, R, c->cascades, c->grid_size, c->scale, c->exp_step_factor, c->max_samples, c->rays_a,
                                    c->t_scratch, c->counter, c->xyzs, c->dirs, c->deltas, c->ts, c->capacity, stream);
}

namespace arn {
thread_local int g_fork_stage = -1;
thread_local cudaEvent_t g_fork_event = nullptr;
thread_local int g_join_stage = -1;
thread_local cudaEvent_t g_join_event = nullptr;
int train_fork(int stage, cudaStream_t st) {
    if (g_join_event && g_join_stage == stage) ARN_CUDA(cudaStreamWaitEvent(st, g_join_event, 0));
    if (g_fork_event && g_fork_stage == stage) ARN_CUDA(cudaEventRecord(g_fork_event, st));
    return ARN_OK;
}
}  // namespace arn
namespace arn {
thread_local LevelGroups g_level_groups = {0, {0}, {nullptr}};
const LevelGroups& level_groups() { return g_level_groups; }
}  // namespace arn
extern "C" ARN_API int arn_train_set_level_groups(int n_groups, const int* level_begin_host, void* const* cuda_events_host) {
    ARN_REQUIRE(n_groups >= 0 && n_groups <= ARN_N_LEVELS, "0..16 groups");
    if (n_groups == 0) { arn::g_level_groups.n = 0; return ARN_OK; }
    ARN_REQUIRE(level_begin_host, "null pointer");
    ARN_REQUIRE(level_begin_host[0] == 0 && level_begin_host[n_groups] == ARN_N_LEVELS, "the groups must cover levels 0..16");
    for (int g = 0; g < n_groups; g++) ARN_REQUIRE(level_begin_host[g] < level_begin_host[g + 1], "level ranges must be increasing and non-empty");
    arn::g_level_groups.n = n_groups;
    for (int g = 0; g <= n_groups; g++) arn::g_level_groups.begin[g] = level_begin_host[g];
    for (int g = 0; g < n_groups; g++) arn::g_level_groups.events[g] = cuda_events_host ? cuda_events_host[g] : nullptr;
    return ARN_OK;
}
extern "C" ARN_API int arn_train_set_join(int stage, void* cuda_event) {
    ARN_REQUIRE(!cuda_event || (stage >= 0 && stage <= 4), "stage must be 0..4");
    arn::g_join_stage = cuda_event ? stage : -1;
    arn::g_join_event = (cudaEvent_t)cuda_event;
    return ARN_OK;
}
extern "C" ARN_API int arn_train_set_fork(int stage, void* cuda_event) {
    ARN_REQUIRE(!cuda_event || (stage >= 0 && stage <= 4), "stage must be 0..4");
    arn::g_fork_stage = cuda_event ? stage : -1;
    arn::g_fork_event = (cudaEvent_t)cuda_event;
    return ARN_OK;
}

// The rest of the step on samples arn_train_march has produced (rays_a, counter, xyzs, dirs, deltas, ts of the config).
extern "C" ARN_API int arn_train_fwbw_marched(const arn_train_t* c, arn_stream_t stream) {
    ARN_REQUIRE(c, "null config");
    ARN_REQUIRE(c->n_rays > 0 && c->capacity > 0, "bad sizes");
    const int64_t R = c->n_rays;
    cudaStream_t st = (cudaStream_t)stream;
    if (int e = train_fork(0, st)) return e;
    if (int e = arn_field_fw_tc_dyn(c->xyzs, c->dirs, c->capacity, c->counter, c->xyz_min_host, c->xyz_max_host, c->levels, c->params_xyz_f16,
                                    c->params_rgb_f16, c->rgb_act, c->ws, c->sigmas, c->rgbs, stream)) return e;
    if (int e = train_fork(1, st)) return e;
    if (int e = arn_composite_train_fw_loss_ex(c->sigmas, c->rgbs, c->deltas, c->ts, c->rays_a, R, c->capacity, c->T_threshold, c->total_samples, c->opacity,
                                               c->depth, c->rgb, c->ws_out, c->rgb_target, c->bg_host, c->lambda_opacity, c->lambda_depth, c->scale,
                                               c->grad_scale, c->rgb_final, c->dL_drgb, c->dL_dopacity, c->dL_ddepth, c->loss_out, /*zero_loss=*/0,
                                               c->dL_dsigmas, c->dL_drgbs, stream)) return e;  // forward + NeRFLoss + backward of every ray in one launch
    if (int e = train_fork(2, st)) return e;
    if (int e = field_bw_tc_impl(c->xyzs, c->capacity, c->counter, c->xyz_min_host, c->xyz_max_host, c->levels, c->params_xyz_f16, c->params_rgb_f16,
                                 c->rgb_act, c->ws, c->sigmas, c->rgbs, c->dL_dsigmas, c->dL_drgbs, c->loss_scale, c->dfeat, c->grad_xyz, c->grad_rgb,
                                 nullptr, /*pack_weights=*/false, stream)) return e;
    return train_fork(4, st);
}

extern "C" ARN_API int arn_train_fwbw(const arn_train_t* c, arn_stream_t stream) {
    if (int e = arn_train_march(c, stream)) return e;
    return arn_train_fwbw_marched(c, stream);
}
