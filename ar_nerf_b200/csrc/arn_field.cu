// libarnerf.so -- the field: multiresolution hash grid + density MLP + SH-4 + colour MLP (replaces tiny-cuda-nn as
// used by models/networks.py:37-78,95-108,133-165).  Numeric contract: fp16 operands, fp32 accumulate.
//
// This file holds the hash-grid encode/backward, SH-4, parameter cast and Adam kernels plus the CUDA-core ("simt")
// MLP kernels.  The simt MLP follows the oracle's operation order exactly (k-ascending fmaf), which makes the whole
// forward bit-comparable with oracle/oracle_field.c; the tensor-core MLP (arn_mlp_tc.cu) is validated against it.
#include "arn_common.cuh"
#include <atomic>
#include "arn_field.cuh"
#include "arn_tc.cuh"

namespace arn {

// ------------------------------------------------------------------------------------------------ parameter cast
__global__ void __launch_bounds__(256) cast_f32_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n && ((uintptr_t)(src + i) & 15) == 0 && ((uintptr_t)(dst + i) & 7) == 0) {
        const float4 v = *reinterpret_cast<const float4*>(src + i);
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dst + i) = o;
    } else {
        for (int k = 0; k < 4 && i + k < n; k++) dst[i + k] = __float2half_rn(src[i + k]);
    }
}

// ------------------------------------------------------------------------------------------------ hash grid
// Forward -- SURVEY Appendix A.3.  One thread per (sample, group of 8 levels): the normalised position is computed once,
// each level keeps its eight gathers in flight, and the thread owns 32 contiguous bytes of the feature row (two 16-byte
// chunks, one full sector) instead of scattering one 4-byte half2 per thread into a sector of its own.  Consecutive
// lanes are consecutive samples of a ray, so at the coarse levels a warp's gathers fall into a few cache lines.
__global__ void __launch_bounds__(256) hash_encode_fw_kernel(const float* __restrict__ xyzs, int64_t n, const int32_t* __restrict__ n_dev, Aabb box,
                                                             const __grid_constant__ LevelTable tbl,
                                                             const __half2* __restrict__ table, __half2* __restrict__ feat, int img,
                                                             const __half* __restrict__ pack_wd, const __half* __restrict__ pack_wc,
                                                             uint8_t* __restrict__ pack_img, int part, int parts) {
    pdl_enter();
    if (n_dev) n = min(n, (int64_t)*n_dev);
    const int grp = blockIdx.y;  // levels [8*grp, 8*grp + 8)
    if (pack_img && grp == 0) {  // rider: the MLP kernels that follow read the weights as a swizzled operand image
        const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
        if (chunk < kWimgChunks) pack_weight_chunk(chunk, pack_wd, pack_wc, pack_img);
    }
    // tile image: the rows that pad the last 128-row tile are written as zeros (the MLP kernels move whole tiles, and the
    // backward multiplies them by zero gradients: they must be finite)
    int64_t n_rows = img ? ((n + 127) & ~(int64_t)127) : n;
    // one of `parts` consecutive ranges of 128-row tiles (the host pipelines range p+1 under the MLP of range p)
    const int64_t tiles_all = (n + 127) / 128;
    const int64_t row0 = (tiles_all * part / parts) * 128;
    n_rows = min(n_rows, (tiles_all * (part + 1) / parts) * 128);
    for (int64_t i = row0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t out[8];
        if (i < n) {
            float x01[3];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const float num = __fsub_rn(xyzs[3 * i + d], box.mn[d]);
                x01[d] = box.inv[d] != 0.0f ? __fmul_rn(num, box.inv[d]) : __fdiv_rn(num, __fsub_rn(box.mx[d], box.mn[d]));
            }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int l = 8 * grp + k;
                float w[3]; uint32_t g[3];
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const float pos = __fmaf_rn(tbl.scale[l], x01[d], 0.5f);
                    const float fl = floorf(pos);
                    w[d] = __fsub_rn(pos, fl); g[d] = (uint32_t)(int32_t)fl;
                }
                uint32_t idx[8]; float wt[8];
                corner_indices(tbl.mode[l], tbl.size[l], tbl.res[l], g, idx);
                corner_weights(w, wt);
                const __half2* lvl = table + tbl.offset[l];
                __half2 tv[8];
#pragma unroll
                for (int c = 0; c < 8; c++) tv[c] = lvl[idx[c]];  // eight independent gathers in flight
                float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const float2 v = __half22float2(tv[c]);
                    acc0 = __fmaf_rn(wt[c], v.x, acc0); acc1 = __fmaf_rn(wt[c], v.y, acc1);
                }
                const __half2 r = __floats2half2_rn(acc0, acc1);
                out[k] = *reinterpret_cast<const uint32_t*>(&r);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++) out[k] = 0u;
        }
        // a 64-byte feature row = 4 chunks of 4 levels; this thread owns chunks 2*grp and 2*grp+1 (tile image: permuted)
        uint4* row = reinterpret_cast<uint4*>(feat + i * ARN_N_LEVELS);
        const uint32_t c0 = 2 * grp, c1 = 2 * grp + 1;
        row[img ? img_chunk64(i, c0) : c0] = make_uint4(out[0], out[1], out[2], out[3]);
        row[img ? img_chunk64(i, c1) : c1] = make_uint4(out[4], out[5], out[6], out[7]);
    }
}

// dfeat row = 128 bytes = 8 chunks of 2 levels (float2 each); tile image: chunk permuted (arn_field.cuh)
__device__ __forceinline__ uint32_t dfeat_col(int img, int64_t i, int l) {
    return img ? ((img_chunk128(i, (uint32_t)l >> 1) << 1) | ((uint32_t)l & 1u)) : (uint32_t)l;
}
__global__ void __launch_bounds__(256) hash_encode_bw_kernel(const float* __restrict__ xyzs, int64_t n, const int32_t* __restrict__ n_dev, Aabb box,
                                                             const __grid_constant__ LevelTable tbl,
                                                             const float2* __restrict__ dfeat, float2* __restrict__ table_grad, int img) {
    if (n_dev) n = min(n, (int64_t)*n_dev);
    const int l = blockIdx.y;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 d = dfeat[i * ARN_N_LEVELS + dfeat_col(img, i, l)];
    if (d.x == 0.0f && d.y == 0.0f) continue;
    float w[3]; uint32_t g[3];
    level_position(xyzs + 3 * i, box, tbl.scale[l], w, g);
    uint32_t idx[8]; float wt[8];
    corner_indices(tbl.mode[l], tbl.size[l], tbl.res[l], g, idx);
    corner_weights(w, wt);
    float2* lvl = table_grad + tbl.offset[l];
#pragma unroll
    for (int c = 0; c < 8; c++) atomicAdd(lvl + idx[c], make_float2(wt[c] * d.x, wt[c] * d.y));
    }
}

// dL/dxyz through the trilinear weights (render_surface_normal, rendering.py:301-313): one thread per sample.
__global__ void __launch_bounds__(256) hash_encode_dx_kernel(const float* __restrict__ xyzs, int64_t n, Aabb box,
                                                             const __grid_constant__ LevelTable tbl,
                                                             const __half2* __restrict__ table, const float2* __restrict__ dfeat,
                                                             float* __restrict__ dL_dxyzs, int img) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gx[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < ARN_N_LEVELS; l++) {
        float w[3]; uint32_t g[3];
        level_position(xyzs + 3 * i, box, tbl.scale[l], w, g);
        const uint32_t size = tbl.size[l], res = tbl.res[l], off = tbl.offset[l];
        const float2 d = dfeat[i * ARN_N_LEVELS + dfeat_col(img, i, l)];
        for (int c = 0; c < 8; c++) {
            uint32_t p[3]; corner_weight(c, w, g, p);
            const float2 t = __half22float2(table[off + grid_index(size, res, p)]);
            const float v = t.x * d.x + t.y * d.y;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                float wd = 1.0f;
#pragma unroll
                for (int e = 0; e < 3; e++) if (e != k) wd *= (c & (1 << e)) ? w[e] : 1.0f - w[e];
                gx[k] += ((c & (1 << k)) ? 1.0f : -1.0f) * wd * v * tbl.scale[l];
            }
        }
    }
    // chain through x01 = (x - min) / (max - min)
    dL_dxyzs[3 * i] = gx[0] / (box.mx[0] - box.mn[0]);
    dL_dxyzs[3 * i + 1] = gx[1] / (box.mx[1] - box.mn[1]);
    dL_dxyzs[3 * i + 2] = gx[2] / (box.mx[2] - box.mn[2]);
}

// Hash-grid backward, run-aggregating form.  Samples arrive ordered along their rays, so at the coarse and middle levels
// several consecutive samples fall into the same cell and hit the same 8 table entries.  One thread walks a segment of
// SEG consecutive samples of ONE level, keeps the 8 corner gradients of the current cell in registers and issues the
// red.global.add.v2.f32 only when the cell changes.  lane % 16 = level: a half-warp reads one full 128-byte dfeat row
// per step and the xyz loads are broadcasts.  (Sums are re-associated relative to the per-sample kernel: same tolerance.)
template <int LPG>
__global__ void __launch_bounds__(256) hash_encode_bw_runs_kernel(const float* __restrict__ xyzs, int64_t n, const int32_t* __restrict__ n_dev, Aabb box,
                                                                  const __grid_constant__ LevelTable tbl, const float2* __restrict__ dfeat,
                                                                  float2* __restrict__ table_grad, int level0, int nlevels, int img,
                                                                  int n_main_blocks, int min_run, WgradReduce red, int part, int parts) {
    pdl_enter();
    __shared__ float red_part[8][32];
    if ((int)blockIdx.x >= n_main_blocks) {  // rider blocks: sum of the MLP weight-gradient slabs (independent of the table work)
        wgrad_reduce_block((int)blockIdx.x - n_main_blocks, red.wpart, red.n_slabs, red.with_rgb, red.dWd, red.dWc, red_part);
        return;
    }
    if (n_dev) n = min(n, (int64_t)*n_dev);
    // LPG consecutive lanes share a run of samples and take the levels level0 .. level0 + LPG - 1 (LPG = 16: the whole row;
    // smaller groups when the caller walks the levels group by group so that a finished group's gradient can leave early)
    const int l = level0 + (threadIdx.x & (LPG - 1));
    if (l >= level0 + nlevels || l >= ARN_N_LEVELS) return;
    // The samples of this launch (one of `parts` consecutive ranges of 128-sample tiles) are cut into EQUAL contiguous runs of
    // `run` samples (chosen by the host: launch_hash_bw_runs), one per lane group, all resident at once: every thread carries
    // the same load, and the runs are as long as the machine allows -- the aggregation below feeds on consecutive samples.
    const int64_t tiles_all = (n + 127) / 128;
    const int64_t s0 = (tiles_all * part / parts) * 128, s1 = min(n, (tiles_all * (part + 1) / parts) * 128);
    const int64_t n_tg = ((int64_t)n_main_blocks * blockDim.x) / LPG;
    int64_t per = (s1 - s0 + n_tg - 1) / n_tg;
    if (per < min_run) per = min_run;
    const int64_t tg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPG;
    const int64_t i0 = s0 + tg * per, i1 = min(s1, i0 + per);
    if (i0 >= i1) return;
    const uint32_t size = tbl.size[l], res = tbl.res[l], mode = tbl.mode[l];
    float2* lvl = table_grad + tbl.offset[l];
    const bool pair_ok = (reinterpret_cast<uintptr_t>(lvl) & 15) == 0;  // 16-byte reductions need the level base aligned
    const float scale = tbl.scale[l];
    uint32_t cg[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu};
    float2 acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = make_float2(0.f, 0.f);
    bool dirty = false;
    // The two x-corners of a cell are neighbouring table entries whenever they form an aligned pair (hashed level: x even,
    // since the hash is x ^ f(y,z); dense level: even linear index): one 16-byte reduction then replaces two 8-byte
    // ones -- the kernel is bound by the number of reduction sectors the L2 can retire.
    auto flush = [&]() {
        uint32_t idx[8];
        corner_indices(mode, size, res, cg, idx);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t i0 = idx[2 * q], i1 = idx[2 * q + 1];
            if (pair_ok && (i0 ^ i1) == 1u) {
                const bool lo0 = (i0 & 1u) == 0u;
                const float2 a = lo0 ? acc[2 * q] : acc[2 * q + 1], b = lo0 ? acc[2 * q + 1] : acc[2 * q];
                atomicAdd(reinterpret_cast<float4*>(lvl + (i0 & ~1u)), make_float4(a.x, a.y, b.x, b.y));
            } else {
                atomicAdd(lvl + i0, acc[2 * q]); atomicAdd(lvl + i1, acc[2 * q + 1]);
            }
            acc[2 * q] = make_float2(0.f, 0.f); acc[2 * q + 1] = make_float2(0.f, 0.f);
        }
    };
    // the walk is a chain of dependent steps (load the sample, locate its cell, maybe flush): the loads of four samples are in
    // flight at a time, so a launch over few levels -- few threads per run -- does not pay the full memory latency per sample
    constexpr int kAhead = 4;
    for (int64_t ib = i0; ib < i1; ib += kAhead) {
        float2 dq[kAhead]; float xq[kAhead][3];
#pragma unroll
        for (int k = 0; k < kAhead; k++) {
            const int64_t i = min(ib + k, i1 - 1);
            dq[k] = dfeat[i * ARN_N_LEVELS + dfeat_col(img, i, l)];
            xq[k][0] = xyzs[3 * i]; xq[k][1] = xyzs[3 * i + 1]; xq[k][2] = xyzs[3 * i + 2];
        }
#pragma unroll
        for (int k = 0; k < kAhead; k++) {
            if (ib + k >= i1) break;
            const float2 d = dq[k];
            float w[3]; uint32_t g[3];
            level_position(xq[k], box, scale, w, g);
            if (g[0] != cg[0] || g[1] != cg[1] || g[2] != cg[2]) {
                if (dirty) { flush(); dirty = false; }
                cg[0] = g[0]; cg[1] = g[1]; cg[2] = g[2];
            }
            if (d.x != 0.0f || d.y != 0.0f) {
                float wt[8];
                corner_weights(w, wt);
#pragma unroll
                for (int c = 0; c < 8; c++) { acc[c].x += wt[c] * d.x; acc[c].y += wt[c] * d.y; }
                dirty = true;
            }
        }
    }
    if (dirty) flush();
}

// Hash-grid backward, warp-segmented form (the default).  A warp takes 32 CONSECUTIVE samples and four levels (one 32-byte
// sector of every sample's dfeat row): each lane owns one sample -- its loads are independent, there is no sequential walk
// -- and, per level, lanes whose samples fall into the same cell form a segment (samples arrive ordered along their rays, so
// equal cells are neighbours): a segmented inclusive scan over the 8 corner sums (shuffles; a doubling step is skipped as
// soon as no segment is longer than its distance) leaves a segment's total in its last lane, which issues the reductions
// -- aligned x-corner pairs as one red.global.add.v4.f32.  At the fine levels every lane is its own segment and the scan is
// skipped altogether.  Unlike the run-walking kernel below, whose time is the length of the walk (a launch over ONE level
// costs 30 us of dependent loads), this kernel's time is its reductions: the levels can be walked group by group
// (arn_train_set_level_groups) at no extra cost, so a finished group's gradient can leave while the next one is reduced.
__global__ void __launch_bounds__(256) hash_encode_bw_warp_kernel(const float* __restrict__ xyzs, int64_t n, const int32_t* __restrict__ n_dev, Aabb box,
                                                                  const __grid_constant__ LevelTable tbl, const float* __restrict__ dfeat,
                                                                  float2* __restrict__ table_grad, int level0, int nlevels, int img,
                                                                  int n_main_blocks, WgradReduce red, int part, int parts) {
    pdl_enter();
    __shared__ float red_part[8][32];
    if ((int)blockIdx.x >= n_main_blocks) {  // rider blocks: sum of the MLP weight-gradient slabs (independent of the table work)
        wgrad_reduce_block((int)blockIdx.x - n_main_blocks, red.wpart, red.n_slabs, red.with_rgb, red.dWd, red.dWc, red_part);
        return;
    }
    if (n_dev) n = min(n, (int64_t)*n_dev);
    const int lane = threadIdx.x & 31;
    const int q0 = level0 >> 2, nq = ((level0 + nlevels + 3) >> 2) - q0;  // level quads touched by [level0, level0 + nlevels)
    const int64_t tiles_all = (n + 127) / 128;
    const int64_t s0 = (tiles_all * part / parts) * 128, s1 = min(n, (tiles_all * (part + 1) / parts) * 128);
    const int64_t n_items = ((s1 - s0 + 31) / 32) * nq;
    const int64_t n_warps = ((int64_t)n_main_blocks * blockDim.x) >> 5;
    for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n_items; item += n_warps) {
        const int64_t i = s0 + (item / nq) * 32 + lane;
        const int q = q0 + (int)(item % nq);
        const bool live = i < s1;
        float x01[3] = {0.f, 0.f, 0.f};
        float4 da = make_float4(0.f, 0.f, 0.f, 0.f), db = da;
        if (live) {
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const float num = __fsub_rn(xyzs[3 * i + d], box.mn[d]);
                x01[d] = box.inv[d] != 0.0f ? __fmul_rn(num, box.inv[d]) : __fdiv_rn(num, __fsub_rn(box.mx[d], box.mn[d]));
            }
            // levels 4q .. 4q+3 = 16-byte chunks 2q, 2q+1 of the 128-byte row (tile image: permuted inside the row, still one sector)
            const float4* row = reinterpret_cast<const float4*>(dfeat + i * (2 * ARN_N_LEVELS));
            da = row[img ? img_chunk128(i, 2u * q) : 2u * q];
            db = row[img ? img_chunk128(i, 2u * q + 1u) : 2u * q + 1u];
        }
        const float2 dl[4] = {make_float2(da.x, da.y), make_float2(da.z, da.w), make_float2(db.x, db.y), make_float2(db.z, db.w)};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int l = 4 * q + k;
            if (l < level0 || l >= level0 + nlevels) continue;  // warp-uniform
            const float2 d = dl[k];
            uint32_t g[3]; float w[3];
#pragma unroll
            for (int e = 0; e < 3; e++) {
                const float pos = __fmaf_rn(tbl.scale[l], x01[e], 0.5f);
                const float fl = floorf(pos);
                w[e] = __fsub_rn(pos, fl); g[e] = (uint32_t)(int32_t)fl;
            }
            const bool nzero = live && (d.x != 0.0f || d.y != 0.0f);
            const unsigned any_nz = __ballot_sync(kFull, nzero);
            if (any_nz == 0u) continue;  // terminated rays: nothing to add for these 32 samples
            float2 acc[8];
            {
                float wt[8];
                corner_weights(w, wt);
#pragma unroll
                for (int c = 0; c < 8; c++) acc[c] = nzero ? make_float2(wt[c] * d.x, wt[c] * d.y) : make_float2(0.f, 0.f);
            }
            // segments of equal cells among neighbouring lanes (a dead lane is its own, empty segment)
            const uint32_t gp0 = __shfl_up_sync(kFull, g[0], 1), gp1 = __shfl_up_sync(kFull, g[1], 1), gp2 = __shfl_up_sync(kFull, g[2], 1);
            const unsigned livem = __ballot_sync(kFull, live);
            const bool head = lane == 0 || !live || !((livem >> (lane - 1)) & 1u) || g[0] != gp0 || g[1] != gp1 || g[2] != gp2;
            const unsigned heads = __ballot_sync(kFull, head);
            if (heads != kFull) {
                const int start = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));  // first lane of my segment
#pragma unroll
                for (int dist = 1; dist < 32; dist <<= 1) {
                    const bool take = lane - dist >= start;
                    if (__ballot_sync(kFull, take) == 0u) break;  // no segment is longer than `dist`
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const float ox = __shfl_up_sync(kFull, acc[c].x, dist), oy = __shfl_up_sync(kFull, acc[c].y, dist);
                        if (take) { acc[c].x += ox; acc[c].y += oy; }
                    }
                }
            }
            // the last lane of a segment holds its total
            const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
            const int start = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));
            const unsigned seg_mask = (0xffffffffu >> (31 - lane)) & (0xffffffffu << start);
            if (tail && (any_nz & seg_mask)) {
                const uint32_t size = tbl.size[l], res = tbl.res[l], mode = tbl.mode[l];
                float2* lvl = table_grad + tbl.offset[l];
                const bool pair_ok = (reinterpret_cast<uintptr_t>(lvl) & 15) == 0;
                uint32_t idx[8];
                corner_indices(mode, size, res, g, idx);
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const uint32_t a0 = idx[2 * p], a1 = idx[2 * p + 1];
                    if (pair_ok && (a0 ^ a1) == 1u) {
                        const bool lo0 = (a0 & 1u) == 0u;
                        const float2 a = lo0 ? acc[2 * p] : acc[2 * p + 1], b = lo0 ? acc[2 * p + 1] : acc[2 * p];
                        atomicAdd(reinterpret_cast<float4*>(lvl + (a0 & ~1u)), make_float4(a.x, a.y, b.x, b.y));
                    } else {
                        atomicAdd(lvl + a0, acc[2 * p]); atomicAdd(lvl + a1, acc[2 * p + 1]);
                    }
                }
            }
        }
    }
}

static int launch_hash_bw_warp(const float* xyzs, int64_t n, const int32_t* n_dev, const Aabb& b, const LevelTable& t, const float* dfeat,
                               float* table_grad, int level0, int nlevels, int img, cudaStream_t st, WgradReduce red, int part, int parts) {
    static int slots = 0;
    if (!slots) {
        int dev = 0, n_sm = 0, per_sm = 0;
        ARN_CUDA(cudaGetDevice(&dev)); ARN_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        ARN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hash_encode_bw_warp_kernel, 256, 0));
        slots = n_sm * (per_sm > 0 ? per_sm : 1);
    }
    const int nq = ((level0 + nlevels + 3) >> 2) - (level0 >> 2);
    const int64_t warps = ((n + 31) / 32) * nq / parts + 1;
    const int cap = tunable(kTunHashBwBlocks) > 0 ? min(slots, 148 * tunable(kTunHashBwBlocks)) : slots;
    const int grid = (int)max((int64_t)1, min((int64_t)cap, (warps + 7) / 8));
    const int riders = red.wpart ? kWgradFloats / 32 : 0;
    ARN_LAUNCH_PDL("hash_encode_bw_warp_kernel", st, (hash_encode_bw_warp_kernel), grid + riders, 256, 0, xyzs, n, n_dev, b, t, dfeat, (float2*)table_grad, level0, nlevels, img, grid, red, part, parts);
    return check_launch("hash_encode_bw_warp");
}

constexpr int kFineLevel0 = 11;  // first level whose cells (NGP geometry, b ~ 1.32-1.66) are crossed in about one marching step
template <int LPG>
static int launch_hash_bw_runs(int min_run, const float* xyzs, int64_t n, const int32_t* n_dev, const Aabb& b, const LevelTable& t, const float* dfeat,
                               float* table_grad, int level0, int nlevels, int img, cudaStream_t st, WgradReduce red, int part, int parts) {
    // one resident wave: blocks per SM from the occupancy calculator (3 at 79 registers), times the SM count
    static int slots = 0;
    if (!slots) {
        int dev = 0, n_sm = 0, per_sm = 0;
        ARN_CUDA(cudaGetDevice(&dev)); ARN_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        ARN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hash_encode_bw_runs_kernel<LPG>, 256, 0));
        slots = n_sm * (per_sm > 0 ? per_sm : 1);
    }
    // Run length: what a full-width launch (16 lanes per run) over the whole machine would give, whatever the group's width --
    // a launch over fewer levels uses proportionally fewer threads instead of shorter runs (shorter runs flush more often,
    // and the kernel is bound by the L2's reduction rate, not by its thread count).
    const int cap = tunable(kTunHashBwBlocks) > 0 ? min(slots, 148 * tunable(kTunHashBwBlocks)) : slots;  // "hash_bw_blocks": blocks per SM (A/B)
    const int64_t n_part = (n + parts - 1) / parts;
    int64_t run = (n_part + ((int64_t)cap * 256 / 16) - 1) / ((int64_t)cap * 256 / 16);
    if (run < min_run) run = min_run;
    // ... except for a group of fine levels only (cells smaller than the sample spacing: nothing to aggregate): short runs, all threads
    if (level0 >= kFineLevel0 && run > min_run) run = max((int64_t)min_run, (n_part * LPG + (int64_t)cap * 256 - 1) / ((int64_t)cap * 256));
    const int64_t threads = ((n_part + run - 1) / run) * LPG;
    const int grid = (int)max((int64_t)1, min((int64_t)cap, (threads + 255) / 256));
    const int riders = red.wpart ? kWgradFloats / 32 : 0;
    ARN_LAUNCH_PDL("hash_encode_bw_runs_kernel", st, (hash_encode_bw_runs_kernel<LPG>), grid + riders, 256, 0, xyzs, n, n_dev, b, t, (const float2*)dfeat, (float2*)table_grad, level0, nlevels, img, grid, min_run, red, part, parts);
    return check_launch("hash_encode_bw_runs");
}
// min_run ("hash_bw_mode", 8..64): the shortest run of consecutive samples one lane group takes (small batches)
static int hash_bw_runs(int min_run, const float* xyzs, int64_t n, const int32_t* n_dev, const Aabb& b, const LevelTable& t, const float* dfeat,
                        float* table_grad, int level0, int nlevels, int img, cudaStream_t st, WgradReduce red = WgradReduce{nullptr, 0, 0, nullptr, nullptr},
                        int part = 0, int parts = 1) {
    // lanes per run: the smallest power of two that holds the group's levels
    if (nlevels > 8) return launch_hash_bw_runs<16>(min_run, xyzs, n, n_dev, b, t, dfeat, table_grad, level0, nlevels, img, st, red, part, parts);
    if (nlevels > 4) return launch_hash_bw_runs<8>(min_run, xyzs, n, n_dev, b, t, dfeat, table_grad, level0, nlevels, img, st, red, part, parts);
    if (nlevels > 2) return launch_hash_bw_runs<4>(min_run, xyzs, n, n_dev, b, t, dfeat, table_grad, level0, nlevels, img, st, red, part, parts);
    return launch_hash_bw_runs<2>(min_run, xyzs, n, n_dev, b, t, dfeat, table_grad, level0, nlevels, img, st, red, part, parts);
}

// ------------------------------------------------------------------------------------------------ SH-4
__global__ void __launch_bounds__(256) sh4_kernel(const float* __restrict__ dirs, int64_t n, __half* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float sh[16];
    sh4_eval(dirs + 3 * i, sh);
    __half2* o = reinterpret_cast<__half2*>(out + 16 * i);
#pragma unroll
    for (int k = 0; k < 8; k++) o[k] = __floats2half2_rn(sh[2 * k], sh[2 * k + 1]);
}

// ------------------------------------------------------------------------------------------------ simt MLPs
// Weights live in shared memory TRANSPOSED ([k][j], j contiguous) so a 128-bit load yields 8 output neurons of one
// input; every thread of a warp reads the same address (broadcast, conflict-free).  One thread per sample.
template <int J, int K>
__device__ __forceinline__ void load_wt(const __half* __restrict__ W, __half* sW) {  // W [J][K] -> sW [K][J]
    for (int e = threadIdx.x; e < J * K; e += blockDim.x) { const int j = e / K, k = e % K; sW[k * J + j] = W[e]; }
}
// acc[j] = sum_k W[j][k] x[k], k ascending (same order as the oracle's matvec)
template <int J, int K>
__device__ __forceinline__ void matvec(const __half* sWt, const __half* x, float* acc) {
#pragma unroll
    for (int j = 0; j < J; j++) acc[j] = 0.0f;
#pragma unroll 4
    for (int k = 0; k < K; k++) {
        const float xk = __half2float(x[k]);
        const uint4* row = reinterpret_cast<const uint4*>(sWt + k * J);
#pragma unroll
        for (int j8 = 0; j8 < J / 8; j8++) {
            const uint4 q = row[j8];
            const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float2 wv = __half22float2(h2[u]);
                acc[j8 * 8 + 2 * u] = __fmaf_rn(wv.x, xk, acc[j8 * 8 + 2 * u]);
                acc[j8 * 8 + 2 * u + 1] = __fmaf_rn(wv.y, xk, acc[j8 * 8 + 2 * u + 1]);
            }
        }
    }
}
// dx[k] = sum_j W[j][k] g[j], j ascending; weights in ORIGINAL layout sW [J][K] (k contiguous)
template <int J, int K>
__device__ __forceinline__ void matvec_t(const __half* sW, const __half* g, float* dx) {
#pragma unroll
    for (int k = 0; k < K; k++) dx[k] = 0.0f;
#pragma unroll 4
    for (int j = 0; j < J; j++) {
        const float gj = __half2float(g[j]);
        const uint4* row = reinterpret_cast<const uint4*>(sW + j * K);
#pragma unroll
        for (int k8 = 0; k8 < K / 8; k8++) {
            const uint4 q = row[k8];
            const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float2 wv = __half22float2(h2[u]);
                dx[k8 * 8 + 2 * u] = __fmaf_rn(wv.x, gj, dx[k8 * 8 + 2 * u]);
                dx[k8 * 8 + 2 * u + 1] = __fmaf_rn(wv.y, gj, dx[k8 * 8 + 2 * u + 1]);
            }
        }
    }
}

// Rows of the saved activations are stored as tile images (arn_field.cuh img_chunk64/128): chunk q of row i sits at the
// permuted position.  NV = 32 halves -> 64-byte rows, NV = 64 -> 128-byte rows.  `buf` is the start of the buffer.
template <int NV>
__device__ __forceinline__ uint32_t img_chunk(int64_t row, uint32_t q) { return NV == 32 ? img_chunk64(row, q) : img_chunk128(row, q); }
template <int NV>
__device__ __forceinline__ void load_row(const __half* __restrict__ buf, int64_t row, __half* dst) {
#pragma unroll
    for (int q = 0; q < NV / 8; q++) reinterpret_cast<uint4*>(dst)[q] = reinterpret_cast<const uint4*>(buf + NV * row)[img_chunk<NV>(row, q)];
}
template <int NV>
__device__ __forceinline__ void store_row(__half* __restrict__ buf, int64_t row, const __half* src) {
#pragma unroll
    for (int q = 0; q < NV / 8; q++) reinterpret_cast<uint4*>(buf + NV * row)[img_chunk<NV>(row, q)] = reinterpret_cast<const uint4*>(src)[q];
}
// shared-memory staging rows of the simt backward (plain layout)
template <int NV>
__device__ __forceinline__ void copy_row(__half* dst, const __half* src) {
#pragma unroll
    for (int q = 0; q < NV / 8; q++) reinterpret_cast<uint4*>(dst)[q] = reinterpret_cast<const uint4*>(src)[q];
}

// Density net 32 -> 64 (ReLU) -> 16 ; sigma = exp(h0)
__global__ void __launch_bounds__(128) density_mlp_fw_simt_kernel(const __half* __restrict__ feat, int64_t n, const __half* __restrict__ Wd,
                                                                  __half* __restrict__ hid, float* __restrict__ h, float* __restrict__ sigmas) {
    __shared__ __align__(16) __half sW1[32 * 64];
    __shared__ __align__(16) __half sW2[64 * 16];
    load_wt<64, 32>(Wd, sW1); load_wt<16, 64>(Wd + 2048, sW2);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __align__(16) __half x[32]; load_row<32>(feat, i, x);
    float acc[64];
    matvec<64, 32>(sW1, x, acc);
    __align__(16) __half hv[64];
#pragma unroll
    for (int j = 0; j < 64; j++) hv[j] = __float2half_rn(fmaxf(acc[j], 0.0f));
    store_row<64>(hid, i, hv);
    float o[16];
    matvec<16, 64>(sW2, hv, o);
#pragma unroll
    for (int q = 0; q < 4; q++) reinterpret_cast<float4*>(h + 16 * i)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    sigmas[i] = expf(o[0]);
}

// Colour net 32 -> 64 -> 64 (ReLU) -> 16 ; input [sh16 | fp16(h16)] ; Sigmoid / None on the first 3 outputs
__global__ void __launch_bounds__(128) rgb_mlp_fw_simt_kernel(const float* __restrict__ dirs, const float* __restrict__ h, int64_t n,
                                                              const __half* __restrict__ Wc, int rgb_act, __half* __restrict__ in32,
                                                              __half* __restrict__ hid1, __half* __restrict__ hid2, float* __restrict__ rgbs) {
    __shared__ __align__(16) __half sW1[32 * 64];
    __shared__ __align__(16) __half sW2[64 * 64];
    __shared__ __align__(16) __half sW3[64 * 16];
    load_wt<64, 32>(Wc, sW1); load_wt<64, 64>(Wc + 2048, sW2); load_wt<16, 64>(Wc + 2048 + 4096, sW3);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __align__(16) __half x[32];
    {
        float sh[16]; sh4_eval(dirs + 3 * i, sh);
#pragma unroll
        for (int k = 0; k < 16; k++) { x[k] = __float2half_rn(sh[k]); x[16 + k] = __float2half_rn(h[16 * i + k]); }
    }
    store_row<32>(in32, i, x);
    float acc[64];
    __align__(16) __half a1[64], a2[64];
    matvec<64, 32>(sW1, x, acc);
#pragma unroll
    for (int j = 0; j < 64; j++) a1[j] = __float2half_rn(fmaxf(acc[j], 0.0f));
    store_row<64>(hid1, i, a1);
    matvec<64, 64>(sW2, a1, acc);
#pragma unroll
    for (int j = 0; j < 64; j++) a2[j] = __float2half_rn(fmaxf(acc[j], 0.0f));
    store_row<64>(hid2, i, a2);
    float o[16];
    matvec<16, 64>(sW3, a2, o);
#pragma unroll
    for (int j = 0; j < 3; j++) rgbs[3 * i + j] = rgb_act ? 1.0f / (1.0f + expf(-o[j])) : o[j];
}

// Backward of both nets (oracle_field.c orc_field_mlp_bw).  Persistent CTAs of 128 threads walk 128-sample tiles.
// Per tile: each thread runs its sample's dgrad chain (bit-identical order with the oracle), stages (g, x) pairs of
// every layer in shared memory, then the CTA accumulates its share of the five weight gradients in registers;
// one atomicAdd per weight per CTA at the end.
template <int J, int K>
__device__ __forceinline__ void wgrad_tile(const __half* sG, const __half* sX, int rows, float* acc) {
    // thread t owns elements e = t + 128*i  (j = e / K, k = e % K); sG [128][64], sX [128][64]
    constexpr int PER = J * K / 128;
    for (int s = 0; s < rows; s++) {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int e = threadIdx.x + 128 * i; const int j = e / K, k = e % K;
            acc[i] = __fmaf_rn(__half2float(sG[s * 64 + j]), __half2float(sX[s * 64 + k]), acc[i]);
        }
    }
}
template <int J, int K>
__device__ __forceinline__ void wgrad_flush(const float* acc, float scale, float* __restrict__ dW) {
    constexpr int PER = J * K / 128;
#pragma unroll
    for (int i = 0; i < PER; i++) atomicAdd(dW + threadIdx.x + 128 * i, acc[i] * scale);
}

__global__ void __launch_bounds__(128) field_mlp_bw_simt_kernel(int64_t n, const float* __restrict__ dL_dsigmas, const float* __restrict__ dL_drgbs,
                                                                const float* __restrict__ rgbs, const float* __restrict__ h,
                                                                const __half* __restrict__ feat, const __half* __restrict__ hid,
                                                                const __half* __restrict__ in32, const __half* __restrict__ hid1,
                                                                const __half* __restrict__ hid2, const __half* __restrict__ Wd,
                                                                const __half* __restrict__ Wc, int rgb_act, float loss_scale,
                                                                float* __restrict__ dWd, float* __restrict__ dWc, float* __restrict__ dfeat) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* sWc1 = reinterpret_cast<__half*>(smem_raw);  // [64][32]
    __half* sWc2 = sWc1 + 2048;                          // [64][64]
    __half* sWc3 = sWc2 + 4096;                          // [16][64]
    __half* sWd1 = sWc3 + 1024;                          // [64][32]
    __half* sWd2 = sWd1 + 2048;                          // [16][64]
    __half* sG = sWd2 + 1024;                            // [128][64]
    __half* sX = sG + 128 * 64;                          // [128][64]
    const bool has_rgb = Wc != nullptr;
    for (int e = threadIdx.x; e < 3072; e += 128) sWd1[e] = Wd[e];  // sWd1|sWd2 contiguous, same layout as params
    if (has_rgb) for (int e = threadIdx.x; e < 7168; e += 128) sWc1[e] = Wc[e];
    __syncthreads();
    float aC1[16], aC2[32], aC3[8], aD1[16], aD2[8];
#pragma unroll
    for (int i = 0; i < 16; i++) { aC1[i] = 0.f; aD1[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 32; i++) aC2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { aC3[i] = 0.f; aD2[i] = 0.f; }
    const float inv_scale = 1.0f / loss_scale;
    const int64_t n_tiles = (n + 127) / 128;
    __half* myG = sG + threadIdx.x * 64; __half* myX = sX + threadIdx.x * 64;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * 128 + threadIdx.x;
        const bool valid = i < n;
        const int rows = (int)min((int64_t)128, n - tile * 128);
        float t[64];
        __align__(16) __half g64[64];
        // ---- colour branch
        if (has_rgb) {
            // output layer
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float g = 0.0f;
                if (valid && j < 3 && dL_drgbs) {
                    const float y = rgbs[3 * i + j];
                    g = dL_drgbs[3 * i + j] * (rgb_act ? y * (1.0f - y) : 1.0f);
                }
                g64[j] = __float2half_rn(g * loss_scale);
            }
            if (valid) { copy_row<16>(myG, g64); load_row<64>(hid2, i, myX); }
            __syncthreads();
            wgrad_tile<16, 64>(sG, sX, rows, aC3);
            __syncthreads();
            if (valid) {
                matvec_t<16, 64>(sWc3, g64, t);
#pragma unroll
                for (int k = 0; k < 64; k++) g64[k] = __float2half_rn(__half2float(myX[k]) > 0.0f ? t[k] : 0.0f);
            }
            __syncthreads();
            if (valid) { copy_row<64>(myG, g64); load_row<64>(hid1, i, myX); }
            __syncthreads();
            wgrad_tile<64, 64>(sG, sX, rows, aC2);
            __syncthreads();
            if (valid) {
                matvec_t<64, 64>(sWc2, g64, t);
#pragma unroll
                for (int k = 0; k < 64; k++) g64[k] = __float2half_rn(__half2float(myX[k]) > 0.0f ? t[k] : 0.0f);
            }
            __syncthreads();
            if (valid) { copy_row<64>(myG, g64); load_row<32>(in32, i, myX); }
            __syncthreads();
            wgrad_tile<64, 32>(sG, sX, rows, aC1);
            __syncthreads();
            if (valid) matvec_t<64, 32>(sWc1, g64, t);  // t[16..31] = scaled dL/dh from the colour branch
        } else {
#pragma unroll
            for (int k = 0; k < 32; k++) t[k] = 0.0f;
        }
        // ---- density branch: dL/dh0 += dL/dsigma * exp(clamp(h0,-15,15))   (custom_functions.py:170-173)
        if (valid) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float g = t[16 + j];
                if (j == 0 && dL_dsigmas) g += (dL_dsigmas[i] * expf(fminf(fmaxf(h[16 * i], -15.0f), 15.0f))) * loss_scale;
                g64[j] = __float2half_rn(g);
            }
            copy_row<16>(myG, g64); load_row<64>(hid, i, myX);
        }
        __syncthreads();
        wgrad_tile<16, 64>(sG, sX, rows, aD2);
        __syncthreads();
        if (valid) {
            matvec_t<16, 64>(sWd2, g64, t);
#pragma unroll
            for (int k = 0; k < 64; k++) g64[k] = __float2half_rn(__half2float(myX[k]) > 0.0f ? t[k] : 0.0f);
        }
        __syncthreads();
        if (valid) { copy_row<64>(myG, g64); load_row<32>(feat, i, myX); }
        __syncthreads();
        wgrad_tile<64, 32>(sG, sX, rows, aD1);
        __syncthreads();
        if (valid) {
            matvec_t<64, 32>(sWd1, g64, t);
#pragma unroll
            for (int q = 0; q < 8; q++)
                reinterpret_cast<float4*>(dfeat + 32 * i)[img_chunk128(i, q)] =
                    make_float4(t[4 * q] * inv_scale, t[4 * q + 1] * inv_scale, t[4 * q + 2] * inv_scale, t[4 * q + 3] * inv_scale);
        }
    }
    if (has_rgb) {
        wgrad_flush<64, 32>(aC1, inv_scale, dWc); wgrad_flush<64, 64>(aC2, inv_scale, dWc + 2048);
        wgrad_flush<16, 64>(aC3, inv_scale, dWc + 2048 + 4096);
    }
    wgrad_flush<64, 32>(aD1, inv_scale, dWd); wgrad_flush<16, 64>(aD2, inv_scale, dWd + 2048);
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam / apex FusedAdam (adam_w_mode=False, wd=0) update, fused with grad un-scale, fp16 refresh, zeroing.
__device__ __forceinline__ void adam_one(int64_t i, float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                         __half* __restrict__ p16, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float inv_gs,
                                         int zero_grad) {
    const float gr = g[i] * inv_gs;
    if (zero_grad) g[i] = 0.0f;
    if (gr == 0.0f && m[i] == 0.0f && v[i] == 0.0f) return;  // untouched hash entry: the update is exactly zero
    const float mi = b1 * m[i] + (1.0f - b1) * gr;
    const float vi = b2 * v[i] + (1.0f - b2) * gr * gr;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    const float np = p[i] - (lr / bc1) * (mi / denom);
    p[i] = np;
    if (p16) p16[i] = __float2half_rn(np);
}
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   __half* __restrict__ p16, int64_t n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2_sqrt, float inv_gs, int zero_grad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    adam_one(i, p, g, m, v, p16, lr, b1, b2, eps, bc1, bc2_sqrt, inv_gs, zero_grad);
}

// 128-bit form: one thread updates 2 x 4 consecutive parameters per turn; all eight 128-bit loads (p, g, m, v of both
// groups) are issued before the first use, so a warp keeps 4 KB in flight (the scalar kernel is latency-bound: two
// dependent DRAM round trips per element).  m, v and the fp32 master are streamed (read once, written once per step:
// ld/st .cs), the gradient and the fp16 copy keep the default policy -- the hash kernels of the next step hit them in L2.
// Same arithmetic per element as adam_kernel.  Requires 16-byte aligned pointers (8 for p16) and n % 4 == 0.
__device__ __forceinline__ void adam_update4(float4& p4, const float4& g4, float4& m4, float4& v4, float lr_bc1, float b1, float b2, float eps,
                                             float bc2_sqrt, float inv_gs) {
    float pn[4] = {p4.x, p4.y, p4.z, p4.w}, mn[4] = {m4.x, m4.y, m4.z, m4.w}, vn[4] = {v4.x, v4.y, v4.z, v4.w};
    const float gr[4] = {g4.x * inv_gs, g4.y * inv_gs, g4.z * inv_gs, g4.w * inv_gs};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (gr[k] == 0.0f && mn[k] == 0.0f && vn[k] == 0.0f) continue;  // untouched hash entry: the update is exactly zero
        mn[k] = b1 * mn[k] + (1.0f - b1) * gr[k];
        vn[k] = b2 * vn[k] + (1.0f - b2) * gr[k] * gr[k];
        const float denom = sqrtf(vn[k]) / bc2_sqrt + eps;
        pn[k] = pn[k] - lr_bc1 * (mn[k] / denom);
    }
    p4 = make_float4(pn[0], pn[1], pn[2], pn[3]); m4 = make_float4(mn[0], mn[1], mn[2], mn[3]); v4 = make_float4(vn[0], vn[1], vn[2], vn[3]);
}
__device__ __forceinline__ bool all_zero4(const float4& a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f && a.w == 0.0f; }
__device__ __forceinline__ uint2 pack_half4(const float4& a) {
    const __half2 lo = __floats2half2_rn(a.x, a.y), hi = __floats2half2_rn(a.z, a.w);
    uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
    return o;
}
// A second, small tensor (the colour net's 7168 parameters) rides in the same launch: the blocks behind `n_main_blocks` update
// it element by element with adam_kernel's arithmetic -- one launch less on the step's serial chain.
struct AdamRider { float* p; float* g; float* m; float* v; __half* p16; int64_t n; };
__global__ void __launch_bounds__(256) adam_vec4_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                                                        uint2* __restrict__ p16, int64_t n4, float lr, float b1, float b2, float eps,
                                                        float bc1, float bc2_sqrt, float inv_gs, int zero_grad, int n_main_blocks, AdamRider rd) {
    pdl_enter();
    const float lr_bc1 = lr / bc1;
    if ((int)blockIdx.x >= n_main_blocks) {
        const int64_t i = (int64_t)(blockIdx.x - n_main_blocks) * blockDim.x + threadIdx.x;
        if (i < rd.n) adam_one(i, rd.p, rd.g, rd.m, rd.v, rd.p16, lr, b1, b2, eps, bc1, bc2_sqrt, inv_gs, zero_grad);
        return;
    }
    const int64_t stride = (int64_t)n_main_blocks * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
        const int64_t j = i + stride;
        const bool two = j < n4;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 ga = g[i], ma = __ldcs(m + i), va = __ldcs(v + i), pa = __ldcs(p + i);
        float4 gb = z, mb = z, vb = z, pb = z;
        if (two) { gb = g[j]; mb = __ldcs(m + j); vb = __ldcs(v + j); pb = __ldcs(p + j); }
        if (!(all_zero4(ga) && all_zero4(ma) && all_zero4(va))) {
            if (zero_grad && !all_zero4(ga)) g[i] = z;
            adam_update4(pa, ga, ma, va, lr_bc1, b1, b2, eps, bc2_sqrt, inv_gs);
            __stcs(m + i, ma); __stcs(v + i, va); __stcs(p + i, pa);
            if (p16) p16[i] = pack_half4(pa);
        }
        if (two && !(all_zero4(gb) && all_zero4(mb) && all_zero4(vb))) {
            if (zero_grad && !all_zero4(gb)) g[j] = z;
            adam_update4(pb, gb, mb, vb, lr_bc1, b1, b2, eps, bc2_sqrt, inv_gs);
            __stcs(m + j, mb); __stcs(v + j, vb); __stcs(p + j, pb);
            if (p16) p16[j] = pack_half4(pb);
        }
    }
}

}  // namespace arn

using namespace arn;

// ================================================================================================ C ABI
extern "C" ARN_API int arn_hashgrid_geometry(int n_levels, int base_resolution, float per_level_scale, int log2_hashmap_size,
                                     float* scale_host, uint32_t* res_host, uint32_t* size_host, uint32_t* offset_host) {
    ARN_REQUIRE(n_levels >= 1 && n_levels <= 32 && base_resolution >= 1 && log2_hashmap_size >= 1 && log2_hashmap_size <= 31, "bad configuration");
    ARN_REQUIRE(scale_host && res_host && size_host && offset_host, "null pointer");
    // tiny-cuda-nn grid.h: scale = exp2f(l * log2f(b)) * N_min - 1 ; res = ceilf(scale) + 1 ; entries = min(roundup8(res^3), 2^T)
    const float log2_pls = log2f(per_level_scale);
    uint32_t offset = 0;
    for (int l = 0; l < n_levels; l++) {
        const float scale = exp2f((float)l * log2_pls) * (float)base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1u;
        const uint32_t max_params = 0xffffffffu / 2u;
        uint32_t params = max_params;
        if ((double)res * res * res < (double)max_params) params = res * res * res;
        params = (params + 7u) / 8u * 8u;
        const uint32_t cap = 1u << log2_hashmap_size;
        if (params > cap) params = cap;
        scale_host[l] = scale; res_host[l] = res; size_host[l] = params; offset_host[l] = offset;
        offset += params;
    }
    offset_host[n_levels] = offset;
    return ARN_OK;
}

extern "C" ARN_API int arn_cast_f32_to_f16(const float* src, void* dst_f16, int64_t n, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(src && dst_f16, "null pointer");
    ARN_LAUNCH("cast_f32_f16_kernel", (cudaStream_t)stream, cast_f32_f16_kernel<<<ceil_div((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst_f16, n));
    return check_launch("cast_f32_to_f16");
}

namespace arn {
int make_levels(const arn_levels_t& lv, LevelTable& t) {
    if (!lv.scale_host || !lv.res_host || !lv.size_host || !lv.offset_host) { set_error("levels: null host table"); return ARN_E_INVALID; }
    for (int l = 0; l < ARN_N_LEVELS; l++) {
        t.scale[l] = lv.scale_host[l]; t.res[l] = lv.res_host[l]; t.size[l] = lv.size_host[l]; t.offset[l] = lv.offset_host[l];
        if (t.size[l] == 0) { set_error("levels: empty level"); return ARN_E_INVALID; }
        // run grid_index's stride loop once (uint32 arithmetic, as on the device) to classify the level
        uint32_t stride = 1; int dims = 0;
        for (int d = 0; d < 3; d++) if (stride <= t.size[l]) { stride *= t.res[l]; dims++; }
        const bool hashed = t.size[l] < stride;
        const bool pow2 = (t.size[l] & (t.size[l] - 1u)) == 0;
        const bool fits = (uint64_t)t.res[l] * t.res[l] * t.res[l] <= (uint64_t)t.size[l];
        t.mode[l] = hashed ? (pow2 ? kIdxHashPow2 : kIdxGeneric) : ((dims == 3 && fits) ? kIdxDense : kIdxGeneric);
    }
    return ARN_OK;
}
int make_box(const float* mn, const float* mx, Aabb& b) {
    if (!mn || !mx) { set_error("aabb: null host pointer"); return ARN_E_INVALID; }
    for (int k = 0; k < 3; k++) {
        b.mn[k] = mn[k]; b.mx[k] = mx[k];
        const float ext = mx[k] - mn[k];
        int e = 0;
        b.inv[k] = (ext > 0.0f && frexpf(ext, &e) == 0.5f && e > -100 && e < 100) ? 1.0f / ext : 0.0f;
    }
    return ARN_OK;
}
}  // namespace arn

namespace arn {
// n is the count (n_dev == nullptr) or the capacity with the real count read on the device from *n_dev (fused step).
inline int sample_grid(int64_t n, const int32_t* n_dev) { return n_dev ? (int)min((int64_t)148 * 8, (n + 255) / 256) : ceil_div(n, 256); }
}
int arn::hash_encode_fw_impl(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                             arn_levels_t levels, const void* table_f16, void* feat_f16, int tile_image, arn_stream_t stream,
                             const __half* pack_wd, const __half* pack_wc, uint8_t* pack_img, int part, int parts) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(xyzs && table_f16 && feat_f16, "null pointer");
    LevelTable t; Aabb b;
    if (int e = make_levels(levels, t)) return e;
    if (int e = make_box(xyz_min_host, xyz_max_host, b)) return e;
    ARN_REQUIRE(((uintptr_t)feat_f16 & 15) == 0, "feat must be 16-byte aligned");
    ARN_REQUIRE(parts >= 1 && part >= 0 && part < parts, "bad part");
    dim3 grid(max(max(sample_grid(n, n_dev) / parts, 1), pack_img ? ceil_div(kWimgChunks, 256) : 1), ARN_N_LEVELS / 8);
    ARN_LAUNCH_PDL("hash_encode_fw_kernel", (cudaStream_t)stream, (hash_encode_fw_kernel), grid, 256, 0, xyzs, n, n_dev, b, t, (const __half2*)table_f16, (__half2*)feat_f16, tile_image,
                                                                                                                           pack_wd, pack_wc, pack_img, part, parts);
    return check_launch("hash_encode_fw");
}
extern "C" ARN_API int arn_hash_encode_fw_dyn(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                                      arn_levels_t levels, const void* table_f16, void* feat_f16, arn_stream_t stream) {
    return hash_encode_fw_impl(xyzs, n, n_dev, xyz_min_host, xyz_max_host, levels, table_f16, feat_f16, 0, stream);
}
extern "C" ARN_API int arn_hash_encode_fw(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                                  arn_levels_t levels, const void* table_f16, void* feat_f16, arn_stream_t stream) {
    return hash_encode_fw_impl(xyzs, n, nullptr, xyz_min_host, xyz_max_host, levels, table_f16, feat_f16, 0, stream);
}

extern "C" ARN_API int arn_hash_encode_bw(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                                  arn_levels_t levels, const void* table_f16, const float* dfeat, float* table_grad,
                                  float* dL_dxyzs, arn_stream_t stream) {
    return hash_encode_bw_impl(xyzs, n, nullptr, xyz_min_host, xyz_max_host, levels, table_f16, dfeat, table_grad, dL_dxyzs, 0, stream);
}
extern "C" ARN_API int arn_hash_encode_bw_dyn(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                                      arn_levels_t levels, const void* table_f16, const float* dfeat, float* table_grad,
                                      float* dL_dxyzs, arn_stream_t stream) {
    return hash_encode_bw_impl(xyzs, n, n_dev, xyz_min_host, xyz_max_host, levels, table_f16, dfeat, table_grad, dL_dxyzs, 0, stream);
}
int arn::hash_encode_bw_impl(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                             arn_levels_t levels, const void* table_f16, const float* dfeat, float* table_grad,
                             float* dL_dxyzs, int tile_image, arn_stream_t stream, WgradReduce red, int part, int parts) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(xyzs && dfeat, "null pointer");
    LevelTable t; Aabb b;
    if (int e = make_levels(levels, t)) return e;
    if (int e = make_box(xyz_min_host, xyz_max_host, b)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (table_grad) {
        const int mode = tunable(kTunHashBwMode);  // 0: one reduction per (sample, level, corner); 1: warp-segmented; >= 8: run-walking, shortest run
        if (mode) {
            // arn_train_set_level_groups: the levels are walked group by group, each group one launch, and the caller's event of
            // a group is recorded behind its launch -- the gradient of those levels is final there and may leave (optimizer /
            // multi-GPU exchange on another stream) while the next group is still reducing.  The MLP weight-gradient slab sum
            // rides in the FIRST launch: the small gradients are final with the first event.
            const LevelGroups& lg = level_groups();
            const bool grouped = lg.n > 0 && parts == 1;
            const int ng = grouped ? lg.n : 1;
            const WgradReduce none{nullptr, 0, 0, nullptr, nullptr};
            for (int g = 0; g < ng; g++) {
                const int l0 = grouped ? lg.begin[g] : 0, l1 = grouped ? lg.begin[g + 1] : ARN_N_LEVELS;
                if (mode == 1) {  // warp-segmented (default)
                    if (int e = launch_hash_bw_warp(xyzs, n, n_dev, b, t, dfeat, table_grad, l0, l1 - l0, tile_image, st, g == 0 ? red : none, part, parts)) return e;
                } else if (int e = hash_bw_runs(mode, xyzs, n, n_dev, b, t, dfeat, table_grad, l0, l1 - l0, tile_image, st, g == 0 ? red : none, part, parts)) return e;
                if (grouped && lg.events[g]) ARN_CUDA(cudaEventRecord((cudaEvent_t)lg.events[g], st));
            }
        } else {
            ARN_REQUIRE(parts == 1, "the per-sample backward kernel is not pipelined");
            dim3 grid(sample_grid(n, n_dev), ARN_N_LEVELS);
            ARN_LAUNCH("hash_encode_bw_kernel", st, hash_encode_bw_kernel<<<grid, 256, 0, st>>>(xyzs, n, n_dev, b, t, (const float2*)dfeat, (float2*)table_grad, tile_image));
            if (int e = check_launch("hash_encode_bw")) return e;
        }
    }
    if (dL_dxyzs) {
        ARN_REQUIRE(table_f16 && !n_dev, "dL_dxyzs needs the table and a host-side count");
        ARN_LAUNCH("hash_encode_dx_kernel", st, hash_encode_dx_kernel<<<ceil_div(n, 256), 256, 0, st>>>(xyzs, n, b, t, (const __half2*)table_f16, (const float2*)dfeat, dL_dxyzs, tile_image));
        if (int e = check_launch("hash_encode_dx")) return e;
    }
    return ARN_OK;
}

extern "C" ARN_API int arn_sh4(const float* dirs, int64_t n, void* out_f16, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(dirs && out_f16, "null pointer");
    ARN_LAUNCH("sh4_kernel", (cudaStream_t)stream, sh4_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(dirs, n, (__half*)out_f16));
    return check_launch("sh4");
}

extern "C" ARN_API int arn_field_fw_simt(const float* xyzs, const float* dirs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                                 arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act,
                                 arn_field_ws_t ws, float* sigmas, float* rgbs, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(xyzs && params_xyz_f16 && ws.feat && ws.hid && ws.h && sigmas, "null pointer");
    const bool with_rgb = dirs != nullptr;
    if (with_rgb) ARN_REQUIRE(params_rgb_f16 && ws.in32 && ws.hid1 && ws.hid2 && rgbs, "null pointer (colour branch)");
    cudaStream_t st = (cudaStream_t)stream;
    const __half* pxyz = (const __half*)params_xyz_f16;
    if (int e = hash_encode_fw_impl(xyzs, n, nullptr, xyz_min_host, xyz_max_host, levels, pxyz + ARN_DENSITY_MLP_PARAMS, ws.feat, 1, stream)) return e;
    ARN_LAUNCH("density_mlp_fw_simt_kernel", st, density_mlp_fw_simt_kernel<<<ceil_div(n, 128), 128, 0, st>>>((const __half*)ws.feat, n, pxyz, (__half*)ws.hid, ws.h, sigmas));
    if (int e = check_launch("density_mlp_fw_simt")) return e;
    if (with_rgb) {
        ARN_LAUNCH("rgb_mlp_fw_simt_kernel", st, rgb_mlp_fw_simt_kernel<<<ceil_div(n, 128), 128, 0, st>>>(dirs, ws.h, n, (const __half*)params_rgb_f16, rgb_act, (__half*)ws.in32,
                                                                (__half*)ws.hid1, (__half*)ws.hid2, rgbs));
        if (int e = check_launch("rgb_mlp_fw_simt")) return e;
    }
    return ARN_OK;
}

extern "C" ARN_API int arn_field_bw_simt(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host, arn_levels_t levels,
                                 const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws,
                                 const float* sigmas, const float* rgbs, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale,
                                 float* dfeat_scratch, float* grad_params_xyz, float* grad_params_rgb, float* dL_dxyzs, arn_stream_t stream) {
    (void)sigmas;
    ARN_REQUIRE(n >= 0 && loss_scale > 0, "bad size / loss_scale");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(xyzs && params_xyz_f16 && ws.feat && ws.hid && ws.h && dfeat_scratch && grad_params_xyz, "null pointer");
    const bool with_rgb = params_rgb_f16 != nullptr && dL_drgbs != nullptr;
    if (with_rgb) ARN_REQUIRE(ws.in32 && ws.hid1 && ws.hid2 && rgbs && grad_params_rgb, "null pointer (colour branch)");
    cudaStream_t st = (cudaStream_t)stream;
    const __half* pxyz = (const __half*)params_xyz_f16;
    const int smem = (7168 + 3072 + 2 * 128 * 64) * (int)sizeof(__half);  // 53,248 B
    // per device, once: SM count and the kernel's dynamic shared memory limit (function attributes are per device)
    static std::atomic<int> n_sm_dev[16] = {};
    int dev = 0;
    ARN_CUDA(cudaGetDevice(&dev));
    ARN_REQUIRE(dev >= 0 && dev < 16, "device index out of range");
    int n_sm = n_sm_dev[dev].load(std::memory_order_acquire);
    if (!n_sm) {
        ARN_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        ARN_CUDA(cudaFuncSetAttribute(field_mlp_bw_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        n_sm_dev[dev].store(n_sm, std::memory_order_release);
    }
    const int64_t n_tiles = (n + 127) / 128;
    const int grid = (int)min((int64_t)n_sm * 4, n_tiles);
    ARN_LAUNCH("field_mlp_bw_simt_kernel", (cudaStream_t)stream, field_mlp_bw_simt_kernel<<<grid, 128, smem, st>>>(n, dL_dsigmas, with_rgb ? dL_drgbs : nullptr, rgbs, ws.h, (const __half*)ws.feat,
                                                     (const __half*)ws.hid, (const __half*)ws.in32, (const __half*)ws.hid1,
                                                     (const __half*)ws.hid2, pxyz, with_rgb ? (const __half*)params_rgb_f16 : nullptr,
                                                     rgb_act, loss_scale, grad_params_xyz, grad_params_rgb, dfeat_scratch));
    if (int e = check_launch("field_mlp_bw_simt")) return e;
    return hash_encode_bw_impl(xyzs, n, nullptr, xyz_min_host, xyz_max_host, levels, pxyz + ARN_DENSITY_MLP_PARAMS, dfeat_scratch,
                               grad_params_xyz + ARN_DENSITY_MLP_PARAMS, dL_dxyzs, 1, stream);
}

extern "C" ARN_API int arn_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* dst_f16, int64_t n, float lr,
                             float beta1, float beta2, float eps, int step, float inv_grad_scale, int zero_grad, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0 && step >= 1, "bad size / step");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(params && grads && exp_avg && exp_avg_sq, "null pointer");
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0 && ((uintptr_t)dst_f16 & 7) == 0;
    if (tunable(kTunAdamVec) && aligned && n % 4 == 0 && n >= 4096) {
        const int64_t n4 = n / 4;
        const int grid = (int)min((int64_t)148 * 16, (n4 + 255) / 256);
        ARN_LAUNCH_PDL("adam_vec4_kernel", st, (adam_vec4_kernel), grid, 256, 0, (float4*)params, (float4*)grads, (float4*)exp_avg, (float4*)exp_avg_sq, (uint2*)dst_f16, n4,
                                                                                  lr, beta1, beta2, eps, bc1, bc2_sqrt, inv_grad_scale, zero_grad, grid, AdamRider{});
    } else {
        ARN_LAUNCH("adam_kernel", st, adam_kernel<<<ceil_div(n, 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, (__half*)dst_f16, n, lr, beta1, beta2,
                                                                                     eps, bc1, bc2_sqrt, inv_grad_scale, zero_grad));
    }
    return check_launch("adam_step");
}

// arn_adam_step for a large tensor and a small one (same hyper-parameters, same step) in ONE launch.
extern "C" ARN_API int arn_adam_step2(float* params, float* grads, float* exp_avg, float* exp_avg_sq, void* dst_f16, int64_t n,
                                      float* params2, float* grads2, float* exp_avg2, float* exp_avg_sq2, void* dst2_f16, int64_t n2,
                                      float lr, float beta1, float beta2, float eps, int step, float inv_grad_scale, int zero_grad, arn_stream_t stream) {
    ARN_REQUIRE(n >= 4096 && n % 4 == 0 && n2 > 0 && n2 <= (1 << 20) && step >= 1, "bad sizes (first tensor: >= 4096 elements, multiple of 4; second: at most 2^20)");
    ARN_REQUIRE(params && grads && exp_avg && exp_avg_sq && params2 && grads2 && exp_avg2 && exp_avg_sq2, "null pointer");
    ARN_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0 && ((uintptr_t)dst_f16 & 7) == 0,
                "the first tensor must be 16-byte aligned (8 for the fp16 copy)");
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n4 = n / 4;
    const int grid = (int)min((int64_t)148 * 16, (n4 + 255) / 256);
    const AdamRider rd{params2, grads2, exp_avg2, exp_avg_sq2, (__half*)dst2_f16, n2};
    ARN_LAUNCH_PDL("adam_vec4_kernel", st, (adam_vec4_kernel), grid + ceil_div(n2, 256), 256, 0, (float4*)params, (float4*)grads, (float4*)exp_avg, (float4*)exp_avg_sq,
                                                                                               (uint2*)dst_f16, n4, lr, beta1, beta2, eps, bc1, bc2_sqrt, inv_grad_scale,
                                                                                               zero_grad, grid, rd);
    return check_launch("adam_step2");
}

// Public entry points: the tensor-core kernels (arn_mlp_tc.cu).
extern "C" int arn_field_fw_tc(const float*, const float*, int64_t, const float*, const float*, arn_levels_t, const void*, const void*, int,
                               arn_field_ws_t, float*, float*, arn_stream_t);
extern "C" int arn_field_bw_tc(const float*, int64_t, const float*, const float*, arn_levels_t, const void*, const void*, int, arn_field_ws_t,
                               const float*, const float*, const float*, const float*, float, float*, float*, float*, float*, arn_stream_t);
extern "C" ARN_API int arn_field_fw(const float* xyzs, const float* dirs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                                    arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act,
                                    arn_field_ws_t ws, float* sigmas, float* rgbs, arn_stream_t stream) {
    return arn_field_fw_tc(xyzs, dirs, n, xyz_min_host, xyz_max_host, levels, params_xyz_f16, params_rgb_f16, rgb_act, ws, sigmas, rgbs, stream);
}
extern "C" ARN_API int arn_field_bw(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host, arn_levels_t levels,
                                    const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws,
                                    const float* sigmas, const float* rgbs, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale,
                                    float* dfeat_scratch, float* grad_params_xyz, float* grad_params_rgb, float* dL_dxyzs, arn_stream_t stream) {
    return arn_field_bw_tc(xyzs, n, xyz_min_host, xyz_max_host, levels, params_xyz_f16, params_rgb_f16, rgb_act, ws, sigmas, rgbs,
                           dL_dsigmas, dL_drgbs, loss_scale, dfeat_scratch, grad_params_xyz, grad_params_rgb, dL_dxyzs, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
namespace arn {
// level-range form of the per-(sample, level) kernel, for diagnostics
__global__ void __launch_bounds__(256) hash_encode_bw_range_kernel(const float* __restrict__ xyzs, int64_t n, const int32_t* __restrict__ n_dev, Aabb box,
                                                                   const __grid_constant__ LevelTable tbl, const float2* __restrict__ dfeat,
                                                                   float2* __restrict__ table_grad, int level0) {
    if (n_dev) n = min(n, (int64_t)*n_dev);
    const int l = blockIdx.y + level0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float2 d = dfeat[i * ARN_N_LEVELS + l];
        if (d.x == 0.0f && d.y == 0.0f) continue;
        float w[3]; uint32_t g[3];
        level_position(xyzs + 3 * i, box, tbl.scale[l], w, g);
        const uint32_t size = tbl.size[l], res = tbl.res[l], off = tbl.offset[l];
#pragma unroll
        for (int c = 0; c < 8; c++) {
            uint32_t p[3]; const float wt = corner_weight(c, w, g, p);
            atomicAdd(table_grad + off + grid_index(size, res, p), make_float2(wt * d.x, wt * d.y));
        }
    }
}
}  // namespace arn

// Measured roof of the hash-grid backward (bench.py "l2_reduction"): red.global.add.v4.f32 to pseudo-random 16-byte aligned
// addresses of an `n_floats` buffer (46 MB = the table gradient: L2-resident), `per_thread` reductions per thread, no other
// work -- the rate at which the L2 retires scattered 16-byte reductions.
__global__ void __launch_bounds__(256) l2_red_peak_kernel(float* __restrict__ buf, uint32_t n_vec4, int per_thread) {
    uint32_t x = ((uint32_t)blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    for (int k = 0; k < per_thread; k++) {
        x = x * 1664525u + 1013904223u;  // LCG: a fresh 16-byte slot per reduction, no locality
        const uint32_t slot = (uint32_t)(((uint64_t)(x >> 4) * n_vec4) >> 28);
        atomicAdd(reinterpret_cast<float4*>(buf) + (slot < n_vec4 ? slot : 0u), make_float4(1.f, 1.f, 1.f, 1.f));
    }
}
extern "C" ARN_API int arn_dbg_l2_red_peak(float* buf, int64_t n_floats, int64_t n_reductions, arn_stream_t stream) {
    ARN_REQUIRE(buf && n_floats >= 1024 && n_reductions > 0 && ((uintptr_t)buf & 15) == 0, "bad arguments");
    const int per_thread = 16;
    const int64_t threads = (n_reductions + per_thread - 1) / per_thread;
    ARN_LAUNCH("l2_red_peak_kernel", (cudaStream_t)stream, l2_red_peak_kernel<<<ceil_div(threads, 256), 256, 0, (cudaStream_t)stream>>>(buf, (uint32_t)(n_floats / 4), per_thread));
    return check_launch("l2_red_peak");
}

// Diagnostics entry (tools/): mode 0 = per-(sample,level) kernel on levels [level0, level0+nlevels), mode = 8/16/32/64 = run-aggregating
// kernel with that segment length.
extern "C" ARN_API int arn_dbg_hash_bw(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host, arn_levels_t levels,
                                       const float* dfeat, float* table_grad, int level0, int nlevels, int mode, arn_stream_t stream) {
    LevelTable t; Aabb b;
    if (int e = make_levels(levels, t)) return e;
    if (int e = make_box(xyz_min_host, xyz_max_host, b)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        dim3 grid(ceil_div(n, 256), nlevels);
        ARN_LAUNCH("hash_encode_bw_range_kernel", st, hash_encode_bw_range_kernel<<<grid, 256, 0, st>>>(xyzs, n, nullptr, b, t, (const float2*)dfeat, (float2*)table_grad, level0));
        return check_launch("dbg_hash_bw");
    }
    return hash_bw_runs(mode, xyzs, n, nullptr, b, t, dfeat, table_grad, level0, nlevels, 0, st);
}
