// Device-side pieces shared by the field kernels: level table, hash indexing, trilinear corner weights, SH-4.
// Operation order is the contract that makes features bit-comparable with oracle/oracle_field.c.
#pragma once
#include "arn_common.cuh"

namespace arn {

struct LevelTable {  // 256 B, passed by value (__grid_constant__)
    float scale[ARN_N_LEVELS];
    uint32_t res[ARN_N_LEVELS];
    uint32_t size[ARN_N_LEVELS];
    uint32_t offset[ARN_N_LEVELS];
};
struct Aabb { float mn[3], mx[3]; };

int make_levels(const arn_levels_t& lv, LevelTable& t);
int make_box(const float* mn, const float* mx, Aabb& b);
// arn_field_bw_tc_dyn with the option of reusing the weight image already in ws.wimg (arn_mlp_tc.cu)
int field_bw_tc_impl(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                     arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws,
                     const float* sigmas, const float* rgbs, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale,
                     float* dfeat_scratch, float* grad_params_xyz, float* grad_params_rgb, float* dL_dxyzs, bool pack_weights, arn_stream_t stream);

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
        const float x01 = __fdiv_rn(__fsub_rn(xyz[d], box.mn[d]), __fsub_rn(box.mx[d], box.mn[d]));
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
