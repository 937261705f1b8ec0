// Device-side pieces shared by the field kernels: level table, hash indexing, trilinear corner weights, SH-4.
// Operation order is the contract that makes features bit-comparable with oracle/oracle_field.c.
#pragma once
#include "arn_common.cuh"

namespace arn {

// Index rule of a level, decided once on the host (make_levels) by running tiny-cuda-nn's grid_index stride loop:
//   kIdxDense   all three strides fit: index = x + y*res + z*res^2, reduced mod size only in the (boundary) case index >= size
//   kIdxHashPow2  hashed and size is a power of two: index = hash & (size - 1)
//   kIdxGeneric anything else: the literal rule (grid_index below)
enum : uint32_t { kIdxDense = 0, kIdxHashPow2 = 1, kIdxGeneric = 2 };
struct LevelTable {  // 320 B, passed by value (__grid_constant__)
    float scale[ARN_N_LEVELS];
    uint32_t res[ARN_N_LEVELS];
    uint32_t size[ARN_N_LEVELS];
    uint32_t offset[ARN_N_LEVELS];
    uint32_t mode[ARN_N_LEVELS];
};
// inv[d] = 1/(mx-mn) when that extent is a power of two (every NGP box: 2*scale), else 0: dividing by a power of two and
// multiplying by its reciprocal round the same real number, so the product is bit-identical with the IEEE division.
struct Aabb { float mn[3], mx[3], inv[3]; };

int make_levels(const arn_levels_t& lv, LevelTable& t);
int make_box(const float* mn, const float* mx, Aabb& b);
// Hash-grid encode / backward with the layout of feat / dfeat selectable: tile_image = 0 is the plain row-major layout of
// the public entry points, 1 the chunk-permuted activation image the MLP kernels bulk-copy (img_chunk64/128 below).
// Optional riders: pack_* != NULL makes the first blocks of the forward build the MLP weight image (arn_tc.cuh);
// red.wpart != NULL makes extra blocks of the run-aggregating backward sum the MLP weight-gradient slabs.
struct WgradReduce { const float* wpart; int n_slabs; int with_rgb; float* dWd; float* dWc; };
int hash_encode_fw_impl(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                        arn_levels_t levels, const void* table_f16, void* feat_f16, int tile_image, arn_stream_t stream,
                        const __half* pack_wd = nullptr, const __half* pack_wc = nullptr, uint8_t* pack_img = nullptr,
                        int part = 0, int parts = 1);
int hash_encode_bw_impl(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                        arn_levels_t levels, const void* table_f16, const float* dfeat, float* table_grad, float* dL_dxyzs,
                        int tile_image, arn_stream_t stream, WgradReduce red = WgradReduce{nullptr, 0, 0, nullptr, nullptr},
                        int part = 0, int parts = 1);
// `part`/`parts`: the call covers one of `parts` consecutive ranges of 128-sample tiles (pipelined field evaluation).
// Side stream + event pool of the calling device for that pipelining (arn_core.cu); events are reused call after call.
struct PipeStreams { cudaStream_t side; cudaEvent_t fork, join, ev[8]; };
int pipe_streams(PipeStreams** out);
// arn_field_bw_tc_dyn with the option of reusing the weight image already in ws.wimg (arn_mlp_tc.cu)
int field_bw_tc_impl(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                     arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws,
                     const float* sigmas, const float* rgbs, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale,
                     float* dfeat_scratch, float* grad_params_xyz, float* grad_params_rgb, float* dL_dxyzs, bool pack_weights, arn_stream_t stream);

// arn_train_set_fork (arn_train.cu): records the caller's event on `st` if `stage` is the selected fork point
int train_fork(int stage, cudaStream_t st);
// arn_train_set_level_groups (arn_train.cu): level ranges [begin[g], begin[g+1]) the hash-grid backward walks one launch at a
// time, in this order, with an event recorded behind each (n == 0: one launch over all levels)
struct LevelGroups { int n; int begin[ARN_N_LEVELS + 1]; void* events[ARN_N_LEVELS]; };
const LevelGroups& level_groups();

// Activation images (arn_mlp_tc.cu): position of logical 16-byte chunk c of row `row` inside the row.  Equal to the
// shared-memory swizzle of a 1024-byte aligned tile with 64- / 128-byte rows (tc::swz<64>, tc::swz<128>); depends on
// row % 8 only, so it is the same for the global row index and the row inside its 128-row tile.
__device__ __forceinline__ uint32_t img_chunk64(int64_t row, uint32_t c) { return c ^ ((uint32_t)(row >> 1) & 3u); }
__device__ __forceinline__ uint32_t img_chunk128(int64_t row, uint32_t c) { return c ^ ((uint32_t)row & 7u); }

// tiny-cuda-nn grid.h grid_index
__device__ __forceinline__ uint32_t grid_index(uint32_t hashmap_size, uint32_t res, const uint32_t p[3]) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (stride <= hashmap_size) { index += p[d] * stride; stride *= res; }
    }
    if (hashmap_size < stride) index = (p[0] * 1u) ^ (p[1] * 2654435761u) ^ (p[2] * 805459861u);
    return index % hashmap_size;
}

// x01 = (x - min) / (max - min)  (networks.py:104, IEEE sub/sub/div as torch does it) ; pos = fma(scale, x01, 0.5)
__device__ __forceinline__ void level_position(const float* __restrict__ xyz, const Aabb& box, float scale, float w[3], uint32_t g[3]) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const float num = __fsub_rn(xyz[d], box.mn[d]);
        const float x01 = box.inv[d] != 0.0f ? __fmul_rn(num, box.inv[d]) : __fdiv_rn(num, __fsub_rn(box.mx[d], box.mn[d]));
        const float pos = __fmaf_rn(scale, x01, 0.5f);
        const float fl = floorf(pos);
        w[d] = __fsub_rn(pos, fl); g[d] = (uint32_t)(int32_t)fl;
    }
}

// weight = prod_d (c_d ? w_d : 1 - w_d), multiplied in the order d = 0,1,2 starting from 1.0f
__device__ __forceinline__ float corner_weight(int c, const float w[3], const uint32_t g[3], uint32_t p[3]) {
    float wt = 1.0f;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (c & (1 << d)) { wt = __fmul_rn(wt, w[d]); p[d] = g[d] + 1u; }
        else { wt = __fmul_rn(wt, __fsub_rn(1.0f, w[d])); p[d] = g[d]; }
    }
    return wt;
}

// The 8 corner indices (c = cx + 2*cy + 4*cz, cell corner g + (cx,cy,cz)) and weights of one (sample, level): same
// values as grid_index / corner_weight corner by corner, with the shared sub-terms computed once.
__device__ __forceinline__ void corner_indices(uint32_t mode, uint32_t size, uint32_t res, const uint32_t g[3], uint32_t idx[8]) {
    if (mode == kIdxHashPow2) {
        const uint32_t m = size - 1u;
        const uint32_t hx[2] = {g[0], g[0] + 1u};
        const uint32_t hy0 = g[1] * 2654435761u, hz0 = g[2] * 805459861u;
        const uint32_t hy[2] = {hy0, hy0 + 2654435761u}, hz[2] = {hz0, hz0 + 805459861u};
#pragma unroll
        for (int c = 0; c < 8; c++) idx[c] = (hx[c & 1] ^ hy[(c >> 1) & 1] ^ hz[(c >> 2) & 1]) & m;
    } else if (mode == kIdxDense) {
        const uint32_t sy = res, sz = res * res;
        const uint32_t b = g[0] + g[1] * sy + g[2] * sz;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            uint32_t i = b + (uint32_t)(c & 1) + ((c & 2) ? sy : 0u) + ((c & 4) ? sz : 0u);
            if (i >= size) i %= size;  // only for x01 == 1 or positions outside the box
            idx[c] = i;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const uint32_t p[3] = {g[0] + (uint32_t)(c & 1), g[1] + (uint32_t)((c >> 1) & 1), g[2] + (uint32_t)((c >> 2) & 1)};
            idx[c] = grid_index(size, res, p);
        }
    }
}
// corner_weight's product ((1*ax)*ay)*az with 1*ax == ax: four xy products shared by the two z values
__device__ __forceinline__ void corner_weights(const float w[3], float wt[8]) {
    const float ax[2] = {__fsub_rn(1.0f, w[0]), w[0]}, ay[2] = {__fsub_rn(1.0f, w[1]), w[1]}, az[2] = {__fsub_rn(1.0f, w[2]), w[2]};
#pragma unroll
    for (int c = 0; c < 8; c++) wt[c] = __fmul_rn(__fmul_rn(ax[c & 1], ay[(c >> 1) & 1]), az[(c >> 2) & 1]);
}

// SH degree 4 of d/|d| routed through u = (d^+1)/2 and x = 2u-1 exactly as networks.py:144-145 + tcnn do (Appendix A.4)
__device__ __forceinline__ void sh4_eval(const float* __restrict__ dir, float sh[16]) {
    const float dx = dir[0], dy = dir[1], dz = dir[2];
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float x = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fdiv_rn(dx, nrm), 1.0f), 2.0f), 2.0f), 1.0f);
    const float y = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fdiv_rn(dy, nrm), 1.0f), 2.0f), 2.0f), 1.0f);
    const float z = __fsub_rn(__fmul_rn(__fdiv_rn(__fadd_rn(__fdiv_rn(dz, nrm), 1.0f), 2.0f), 2.0f), 1.0f);
    const float xy = __fmul_rn(x, y), xz = __fmul_rn(x, z), yz = __fmul_rn(y, z);
    const float x2 = __fmul_rn(x, x), y2 = __fmul_rn(y, y), z2 = __fmul_rn(z, z);
    sh[0] = 0.28209479177387814f;
    sh[1] = __fmul_rn(-0.48860251190291987f, y);
    sh[2] = __fmul_rn(0.48860251190291987f, z);
    sh[3] = __fmul_rn(-0.48860251190291987f, x);
    sh[4] = __fmul_rn(1.0925484305920792f, xy);
    sh[5] = __fmul_rn(-1.0925484305920792f, yz);
    sh[6] = __fsub_rn(__fmul_rn(0.94617469575755997f, z2), 0.31539156525251999f);
    sh[7] = __fmul_rn(-1.0925484305920792f, xz);
    sh[8] = __fsub_rn(__fmul_rn(0.54627421529603959f, x2), __fmul_rn(0.54627421529603959f, y2));
    sh[9] = __fmul_rn(__fmul_rn(0.59004358992664352f, y), __fadd_rn(__fmul_rn(-3.0f, x2), y2));
    sh[10] = __fmul_rn(__fmul_rn(2.8906114426405538f, xy), z);
    sh[11] = __fmul_rn(__fmul_rn(0.45704579946446572f, y), __fsub_rn(1.0f, __fmul_rn(5.0f, z2)));
    sh[12] = __fmul_rn(__fmul_rn(0.3731763325901154f, z), __fsub_rn(__fmul_rn(5.0f, z2), 3.0f));
    sh[13] = __fmul_rn(__fmul_rn(0.45704579946446572f, x), __fsub_rn(1.0f, __fmul_rn(5.0f, z2)));
    sh[14] = __fmul_rn(__fmul_rn(1.4453057213202769f, z), __fsub_rn(x2, y2));
    sh[15] = __fmul_rn(__fmul_rn(0.59004358992664352f, x), __fadd_rn(-x2, __fmul_rn(3.0f, y2)));
}

}  // namespace arn
