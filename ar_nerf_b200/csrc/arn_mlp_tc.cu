// libarnerf.so -- tensor-core MLPs of the field (density 32-64-16, colour 32-64-64-16) on tcgen05 / TMEM.
//
// Forward, one CTA of 128 threads per 128-sample tile (persistent over tiles, up to 4 CTAs per SM):
//   * the five weight matrices arrive once per CTA as ONE bulk (TMA) copy of a pre-swizzled 20 KB operand image;
//   * thread t owns sample row t: it stages the row as a swizzled K-major A tile in shared memory, one elected thread
//     issues the layer's tcgen05.mma chain (M=128, N=64|16, K=16 per instruction, fp32 accumulators in TMEM) and commits
//     to an mbarrier, every thread then pulls ITS row of the accumulator with tcgen05.ld (32x32b: lane == row),
//     applies the activation, rounds to fp16 and writes the next layer's A tile (+ the saved activation for backward).
//   Numeric contract: fp16 operands, fp32 accumulate -- identical rounding points to the simt
//   kernels and the oracle; only the accumulation order inside the MMA differs.
#include "arn_common.cuh"
#include "arn_field.cuh"
#include "arn_tc.cuh"

namespace arn {

using namespace tc;

// fp16 params (row-major [out][in] per layer, tcnn order) -> swizzled operand image (arn_tc.cuh)
__global__ void __launch_bounds__(256) pack_mlp_weights_kernel(const __half* __restrict__ Wd, const __half* __restrict__ Wc,
                                                               uint8_t* __restrict__ img) {
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk (8 halves) per thread
    // chunk ranges per layer: D1 256, D2 128, C1 256, C2 512, C3 128  -> 1280 chunks
    const __half* src; uint32_t dst;
    if (chunk < 256) { const int r = chunk >> 2, c = chunk & 3; src = Wd + r * 32 + c * 8; dst = kWimgD1 + swz<64>(r, c); }
    else if (chunk < 384) { const int q = chunk - 256, r = q >> 3, c = q & 7; src = Wd + 2048 + r * 64 + c * 8; dst = kWimgD2 + swz<128>(r, c); }
    else if (chunk < 640) { const int q = chunk - 384, r = q >> 2, c = q & 3; src = Wc ? Wc + r * 32 + c * 8 : nullptr; dst = kWimgC1 + swz<64>(r, c); }
    else if (chunk < 1152) { const int q = chunk - 640, r = q >> 3, c = q & 7; src = Wc ? Wc + 2048 + r * 64 + c * 8 : nullptr; dst = kWimgC2 + swz<128>(r, c); }
    else if (chunk < 1280) { const int q = chunk - 1152, r = q >> 3, c = q & 7; src = Wc ? Wc + 6144 + r * 64 + c * 8 : nullptr; dst = kWimgC3 + swz<128>(r, c); }
    else return;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (src) v = *reinterpret_cast<const uint4*>(src);
    *reinterpret_cast<uint4*>(img + dst) = v;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Pull NQ*16 accumulator columns of this thread's row, apply ReLU, round to fp16: out[NQ*8] packed half2.
template <int NQ, bool RELU>
__device__ __forceinline__ void epilogue_row_f16(uint32_t taddr, uint32_t* out) {
    float v[NQ * 16];
#pragma unroll
    for (int q = 0; q < NQ; q++) tmem_ld16(taddr + 16 * q, v + 16 * q);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < NQ * 8; j++) {
        float a = v[2 * j], b = v[2 * j + 1];
        if (RELU) { a = fmaxf(a, 0.0f); b = fmaxf(b, 0.0f); }
        out[j] = pack2(a, b);
    }
}

constexpr int kFwSmemTile32 = 128 * 64;    // 128 rows x 32 halves
constexpr int kFwSmemTile64 = 128 * 128;   // 128 rows x 64 halves
constexpr int kFwSmemBytes = kWimgBytes + kFwSmemTile32 + kFwSmemTile64 + 1024;  // + alignment slack

__global__ void __launch_bounds__(128) field_mlp_fw_tc_kernel(const __half* __restrict__ feat, const float* __restrict__ dirs, int64_t n,
                                                              const int32_t* __restrict__ n_dev, const uint8_t* __restrict__ wimg, int rgb_act, int with_rgb,
                                                              __half* __restrict__ hid, float* __restrict__ h, float* __restrict__ sigmas,
                                                              __half* __restrict__ in32, __half* __restrict__ hid1, __half* __restrict__ hid2,
                                                              float* __restrict__ rgbs) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    if (n_dev) n = min(n, (int64_t)*n_dev);  // fused step: the sample count lives on the device
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzled tiles need 1024-byte alignment
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sW = base, sT32 = base + kWimgBytes, sT64 = sT32 + kFwSmemTile32;
    uint8_t* pT32 = sm + kWimgBytes; uint8_t* pT64 = pT32 + kFwSmemTile32;
    const uint32_t bar_w = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) { mbar_init(bar_w, 1); mbar_init(bar_mma, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t wbytes = with_rgb ? kWimgBytes : kWimgC1;
    if (tid == 0) { mbar_expect_tx(bar_w, wbytes); bulk_g2s(sW, wimg, wbytes, bar_w); }
    mbar_wait(bar_w, 0);

    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes; lane == sample row
    const uint64_t aT32 = smem_desc<64>(sT32), aT64 = smem_desc<128>(sT64);
    const uint64_t bD1 = smem_desc<64>(sW + kWimgD1), bD2 = smem_desc<128>(sW + kWimgD2);
    const uint64_t bC1 = smem_desc<64>(sW + kWimgC1), bC2 = smem_desc<128>(sW + kWimgC2), bC3 = smem_desc<128>(sW + kWimgC3);
    constexpr uint32_t kI64 = instr_desc(128, 64, 0, 0), kI16 = instr_desc(128, 16, 0, 0);
    uint32_t phase = 0;

    // issue one layer: D[tmem cols] = A (128 x K) * W^T, K = 16 * ksteps; then commit
    auto issue_only = [&](uint32_t dcol, uint64_t a, uint64_t b, uint32_t idesc, int ksteps) {
        fence_before_sync(); fence_async_smem(); __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            for (int k = 0; k < ksteps; k++) mma_f16(tmem + dcol, desc_advance(a, 32 * k), desc_advance(b, 32 * k), idesc, k > 0);
            mma_commit(bar_mma);
        }
    };
    auto wait_mma = [&]() {
        mbar_wait(bar_mma, phase); phase ^= 1;
        fence_after_sync();
    };
    auto issue = [&](uint32_t dcol, uint64_t a, uint64_t b, uint32_t idesc, int ksteps) { issue_only(dcol, a, b, idesc, ksteps); wait_mma(); };

    // The tile's only global inputs (feature row, ray direction) are requested one tile ahead, between the issue of the
    // previous tile's last MMA and its completion (ptxas cannot hoist them above that bar.sync; see the backward kernel).
    const int64_t n_tiles = (n + 127) / 128;
    uint4 fr[4]; float dr[3] = {1.0f, 0.0f, 0.0f};
    auto fetch_inputs = [&](int64_t t) {
        const int64_t i = t * 128 + tid;
        const bool ok = t < n_tiles && i < n;
#pragma unroll
        for (int c = 0; c < 4; c++) fr[c] = ok ? __ldg(reinterpret_cast<const uint4*>(feat + 32 * i) + c) : make_uint4(0, 0, 0, 0);
        if (with_rgb) {
            dr[0] = ok ? dirs[3 * i] : 1.0f; dr[1] = ok ? dirs[3 * i + 1] : 0.0f; dr[2] = ok ? dirs[3 * i + 2] : 0.0f;
        }
    };
    fetch_inputs(blockIdx.x);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * 128 + tid;
        const bool valid = i < n;
        // ---- density layer 1: A = feat row (32 halves)
#pragma unroll
        for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(pT32 + swz<64>(tid, c)) = fr[c];
        const float dcur[3] = {dr[0], dr[1], dr[2]};
        issue_only(0, aT32, bD1, kI64, 2);
        if (!with_rgb) fetch_inputs(tile + gridDim.x);
        wait_mma();
        {
            uint32_t o[32];
            epilogue_row_f16<4, true>(trow + 0, o);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const uint4 v = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                *reinterpret_cast<uint4*>(pT64 + swz<128>(tid, c)) = v;
                if (valid) reinterpret_cast<uint4*>(hid + 64 * i)[c] = v;
            }
        }
        // ---- density layer 2 -> h (16, fp32), sigma
        issue(64, aT64, bD2, kI16, 4);
        {
            float hv[16];
            tmem_ld16(trow + 64, hv); tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4; q++) reinterpret_cast<float4*>(h + 16 * i)[q] = make_float4(hv[4 * q], hv[4 * q + 1], hv[4 * q + 2], hv[4 * q + 3]);
                sigmas[i] = expf(hv[0]);
            }
            if (with_rgb) {  // colour-net input row [sh16 | fp16(h16)]
                float sh[16];
                sh4_eval(dcur, sh);
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 8; j++) { o[j] = pack2(sh[2 * j], sh[2 * j + 1]); o[8 + j] = pack2(hv[2 * j], hv[2 * j + 1]); }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint4 v = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                    *reinterpret_cast<uint4*>(pT32 + swz<64>(tid, c)) = v;
                    if (valid) reinterpret_cast<uint4*>(in32 + 32 * i)[c] = v;
                }
            }
        }
        if (!with_rgb) continue;
        // ---- colour layer 1
        issue(0, aT32, bC1, kI64, 2);
        {
            uint32_t o[32];
            epilogue_row_f16<4, true>(trow + 0, o);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const uint4 v = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                *reinterpret_cast<uint4*>(pT64 + swz<128>(tid, c)) = v;
                if (valid) reinterpret_cast<uint4*>(hid1 + 64 * i)[c] = v;
            }
        }
        // ---- colour layer 2 (its A tile is overwritten by its own output once the MMA has completed)
        issue(64, aT64, bC2, kI64, 4);
        {
            uint32_t o[32];
            epilogue_row_f16<4, true>(trow + 64, o);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const uint4 v = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                *reinterpret_cast<uint4*>(pT64 + swz<128>(tid, c)) = v;
                if (valid) reinterpret_cast<uint4*>(hid2 + 64 * i)[c] = v;
            }
        }
        // ---- colour layer 3 -> rgb
        issue_only(0, aT64, bC3, kI16, 4);
        fetch_inputs(tile + gridDim.x);
        wait_mma();
        {
            float ov[16];
            tmem_ld16(trow + 0, ov); tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 3; j++) rgbs[3 * i + j] = rgb_act ? 1.0f / (1.0f + expf(-ov[j])) : ov[j];
            }
        }
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace arn

using namespace arn;

extern "C" int arn_field_fw_tc_dyn(const float*, const float*, int64_t, const int32_t*, const float*, const float*, arn_levels_t, const void*, const void*,
                                   int, arn_field_ws_t, float*, float*, arn_stream_t);
extern "C" int arn_field_bw_tc_dyn(const float*, int64_t, const int32_t*, const float*, const float*, arn_levels_t, const void*, const void*, int,
                                   arn_field_ws_t, const float*, const float*, const float*, const float*, float, float*, float*, float*, float*, arn_stream_t);
extern "C" int arn_hash_encode_fw_dyn(const float*, int64_t, const int32_t*, const float*, const float*, arn_levels_t, const void*, void*, arn_stream_t);
extern "C" int arn_hash_encode_bw_dyn(const float*, int64_t, const int32_t*, const float*, const float*, arn_levels_t, const void*, const float*, float*,
                                      float*, arn_stream_t);

extern "C" ARN_API int arn_field_fw_tc(const float* xyzs, const float* dirs, int64_t n, const float* xyz_min_host, const float* xyz_max_host,
                                       arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act,
                                       arn_field_ws_t ws, float* sigmas, float* rgbs, arn_stream_t stream) {
    return arn_field_fw_tc_dyn(xyzs, dirs, n, nullptr, xyz_min_host, xyz_max_host, levels, params_xyz_f16, params_rgb_f16, rgb_act, ws, sigmas, rgbs, stream);
}
// n_dev != NULL: n is the capacity of the buffers and the sample count is read on the device (fused training step).
extern "C" ARN_API int arn_field_fw_tc_dyn(const float* xyzs, const float* dirs, int64_t n, const int32_t* n_dev, const float* xyz_min_host,
                                           const float* xyz_max_host, arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16,
                                           int rgb_act, arn_field_ws_t ws, float* sigmas, float* rgbs, arn_stream_t stream) {
    ARN_REQUIRE(n >= 0, "bad size");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(xyzs && params_xyz_f16 && ws.feat && ws.hid && ws.h && ws.wimg && sigmas, "null pointer");
    const bool with_rgb = dirs != nullptr;
    if (with_rgb) ARN_REQUIRE(params_rgb_f16 && ws.in32 && ws.hid1 && ws.hid2 && rgbs, "null pointer (colour branch)");
    cudaStream_t st = (cudaStream_t)stream;
    const __half* pxyz = (const __half*)params_xyz_f16;
    if (int e = arn_hash_encode_fw_dyn(xyzs, n, n_dev, xyz_min_host, xyz_max_host, levels, pxyz + ARN_DENSITY_MLP_PARAMS, ws.feat, stream)) return e;
    ARN_LAUNCH("pack_mlp_weights_kernel", st, pack_mlp_weights_kernel<<<5, 256, 0, st>>>(pxyz, (const __half*)params_rgb_f16, (uint8_t*)ws.wimg));
    if (int e = check_launch("pack_mlp_weights")) return e;
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0; ARN_CUDA(cudaGetDevice(&dev)); ARN_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        ARN_CUDA(cudaFuncSetAttribute(field_mlp_fw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwSmemBytes));
    }
    const int64_t n_tiles = (n + 127) / 128;
    const int grid = (int)(n_tiles < (int64_t)n_sm * 4 ? n_tiles : (int64_t)n_sm * 4);
    ARN_LAUNCH("field_mlp_fw_tc_kernel", st, field_mlp_fw_tc_kernel<<<grid, 128, kFwSmemBytes, st>>>(
        (const __half*)ws.feat, dirs, n, n_dev, (const uint8_t*)ws.wimg, rgb_act, with_rgb ? 1 : 0, (__half*)ws.hid, ws.h, sigmas,
        (__half*)ws.in32, (__half*)ws.hid1, (__half*)ws.hid2, rgbs));
    return check_launch("field_mlp_fw_tc");
}

namespace arn {
using namespace tc;
// =====================================================================================================================
// Backward.  Same CTA shape (128 threads = 128 sample rows, persistent over tiles, 2 CTAs per SM).
// Per layer, walking the net backwards, ONE commit covers two MMA chains that read the same two shared-memory tiles:
//   dgrad   R[128 x in]  = G (K-major A: rows = samples, K = out)  x  W (MN-major B straight from the forward's weight image)
//   wgrad   dW[out x in] += G^T X : A = G as MN-major (M = out), B = X as MN-major (N = in), K = the 128 samples of the tile
// The five weight-gradient accumulators stay in TENSOR MEMORY for the whole kernel (160 columns) and are flushed once
// per CTA with atomicAdd; layers with 16 outputs accumulate the transposed product (M = in = 64, N = 16).
// Thread t then pulls row t of R with tcgen05.ld, applies the ReLU mask of its own activation row, scales/rounds to
// fp16 and writes the next G tile.  Rounding points are those of the simt kernel / oracle (fp16 G, fp32 accumulate).
// TMEM map: [0,64) R | [64,80) dWc3^T | [80,144) dWc2 | [144,176) dWc1 | [176,192) dWd2^T | [192,224) dWd1   (256 allocated)
constexpr int kBwSmemBytes = kWimgBytes + 3 * kFwSmemTile64 + 1024;
constexpr int kWgradFloats = 7168 + 3072;  // one slab of per-CTA weight gradients
constexpr int kMaxBwCtas = ARN_FIELD_SCRATCH_SLABS;
constexpr uint32_t kColR = 0, kColC3 = 64, kColC2 = 80, kColC1 = 144, kColD2 = 176, kColD1 = 192;

template <int RB>
__device__ __forceinline__ void stage_row(uint8_t* tile, int row, const __half* __restrict__ src, bool valid) {
#pragma unroll
    for (int c = 0; c < RB / 16; c++) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (valid) v = reinterpret_cast<const uint4*>(src)[c];
        *reinterpret_cast<uint4*>(tile + swz<RB>(row, c)) = v;
    }
}

// t[NQ*16] = this thread's row of R; g = fp16(relu'(x) * t) written as a RB=128 row of `gtile`; mask from row of `xtile`.
__device__ __forceinline__ void epilogue_mask64(uint32_t taddr, const uint8_t* xtile, uint8_t* gtile, int row) {
    float v[64];
#pragma unroll
    for (int q = 0; q < 4; q++) tmem_ld16(taddr + 16 * q, v + 16 * q);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xtile + swz<128>(row, c));
        const __half2* xh = reinterpret_cast<const __half2*>(&xv);
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float2 xf = __half22float2(xh[u]);
            o[u] = pack2(xf.x > 0.0f ? v[8 * c + 2 * u] : 0.0f, xf.y > 0.0f ? v[8 * c + 2 * u + 1] : 0.0f);
        }
        *reinterpret_cast<uint4*>(gtile + swz<128>(row, c)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void __launch_bounds__(128) field_mlp_bw_tc_kernel(int64_t n, const int32_t* __restrict__ n_dev, const float* __restrict__ dL_dsigmas, const float* __restrict__ dL_drgbs,
                                                              const float* __restrict__ rgbs, const float* __restrict__ h,
                                                              const __half* __restrict__ feat, const __half* __restrict__ hid,
                                                              const __half* __restrict__ in32, const __half* __restrict__ hid1,
                                                              const __half* __restrict__ hid2, const uint8_t* __restrict__ wimg, int rgb_act,
                                                              int with_rgb, float loss_scale, float* __restrict__ dWd, float* __restrict__ dWc,
                                                              float* __restrict__ dfeat, float* __restrict__ wpart) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    if (n_dev) n = min(n, (int64_t)*n_dev);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sW = base, sGa = base + kWimgBytes, sGb = sGa + kFwSmemTile64, sX = sGb + kFwSmemTile64;
    uint8_t* pGa = sm + kWimgBytes; uint8_t* pGb = pGa + kFwSmemTile64; uint8_t* pX = pGb + kFwSmemTile64;
    const uint32_t bar_w = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) { mbar_init(bar_w, 1); mbar_init(bar_mma, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t wbytes = with_rgb ? kWimgBytes : kWimgC1;
    if (tid == 0) { mbar_expect_tx(bar_w, wbytes); bulk_g2s(sW, wimg, wbytes, bar_w); }
    mbar_wait(bar_w, 0);

    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t phase = 0;
    const float inv_scale = 1.0f / loss_scale;
    uint32_t acc = 0;  // 0 on the CTA's first tile: the wgrad accumulators are initialised by the MMA itself

    // One layer step: wgrad (K = 128 samples, 8 MMAs) + dgrad (K = out, ksteps MMAs), one commit.
    //   g   : G tile (rows = samples, RBG bytes per row)        x : X tile (RBX bytes per row)
    //   w   : weight tile of this layer in the image (MN-major B for dgrad), RBW bytes per row (= 2*in)
    //   wgrad M=64: A = (t_out16 ? x : g) MN-major, B = (t_out16 ? g : x) MN-major
    auto layer = [&](uint64_t g_desc, int rbg, uint64_t x_desc, int rbx, uint64_t w_desc, int rbw, int n_in, int n_out, uint32_t wcol) {
        fence_before_sync(); fence_async_smem(); __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const bool t16 = n_out == 16;
            const uint64_t wa = t16 ? x_desc : g_desc, wb = t16 ? g_desc : x_desc;
            const int rba = t16 ? rbx : rbg, rbb = t16 ? rbg : rbx;
            const uint32_t wi = instr_desc(64, t16 ? 16 : n_in, 1, 1);
            for (int k = 0; k < 8; k++)  // 16 samples per MMA = 16 rows of each tile
                mma_f16(tmem + wcol, desc_advance(wa, 16 * rba * k), desc_advance(wb, 16 * rbb * k), wi, acc | (uint32_t)(k > 0));
            const uint32_t di = instr_desc(128, n_in, 0, 1);
            for (int k = 0; k < n_out / 16; k++)  // K = out: 32 B along a G row, 16 rows down the weight tile
                mma_f16(tmem + kColR, desc_advance(g_desc, 32 * k), desc_advance(w_desc, 16 * rbw * k), di, k > 0);
            mma_commit(bar_mma);
        }
    };
    auto layer_wait = [&]() {
        mbar_wait(bar_mma, phase); phase ^= 1;
        fence_after_sync();
    };

    const uint64_t dGa128 = smem_desc<128>(sGa), dGa32 = smem_desc<32>(sGa), dGb128 = smem_desc<128>(sGb), dGb32 = smem_desc<32>(sGb);
    const uint64_t dX128 = smem_desc<128>(sX), dX64 = smem_desc<64>(sX);
    const uint64_t wD1 = smem_desc<64>(sW + kWimgD1), wD2 = smem_desc<128>(sW + kWimgD2);
    const uint64_t wC1 = smem_desc<64>(sW + kWimgC1), wC2 = smem_desc<128>(sW + kWimgC2), wC3 = smem_desc<128>(sW + kWimgC3);

    // Global -> register prefetch: the activation row a layer needs is requested one layer ahead (and the first layer's
    // inputs of the NEXT tile during the last layer), so its DRAM/L2 latency overlaps the MMA round trip and the epilogue
    // instead of being exposed five times per tile.
    // The consumer of a prefetched row must come BEFORE the next prefetch is issued: loads retire through in-order
    // scoreboards, so a use placed after younger loads waits for those as well (seen in ncu as long-scoreboard stalls on
    // the first use).  ptxas hoists loads to the top of a basic block, so every fetch sits between layer() -- whose
    // bar.sync it cannot cross -- and layer_wait(), after the previous row has gone to shared memory.
    auto fetch_row128 = [&](uint4* r, const __half* src, bool ok) {
#pragma unroll
        for (int c = 0; c < 8; c++) r[c] = ok ? __ldg(reinterpret_cast<const uint4*>(src) + c) : make_uint4(0, 0, 0, 0);
    };
    auto fetch_row64 = [&](uint4* r, const __half* src, bool ok) {
#pragma unroll
        for (int c = 0; c < 4; c++) r[c] = ok ? __ldg(reinterpret_cast<const uint4*>(src) + c) : make_uint4(0, 0, 0, 0);
    };
    auto put_row128 = [&](const uint4* r) {
#pragma unroll
        for (int c = 0; c < 8; c++) *reinterpret_cast<uint4*>(pX + swz<128>(tid, c)) = r[c];
    };
    auto put_row64 = [&](const uint4* r) {
#pragma unroll
        for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(pX + swz<64>(tid, c)) = r[c];
    };

    const int64_t n_tiles = (n + 127) / 128;
    uint4 xr[8], xs[4];           // prefetched activation rows (128-byte / 64-byte)
    float pre_y[3], pre_g[3];     // prefetched rgbs / dL_drgbs of this thread's sample
    {
        const int64_t i = (int64_t)blockIdx.x * 128 + tid;
        const bool valid = blockIdx.x < n_tiles && i < n;
        if (with_rgb) {
            fetch_row128(xr, hid2 + 64 * i, valid);
#pragma unroll
            for (int j = 0; j < 3; j++) { pre_y[j] = valid ? rgbs[3 * i + j] : 0.0f; pre_g[j] = (valid && dL_drgbs) ? dL_drgbs[3 * i + j] : 0.0f; }
        } else {
            fetch_row128(xr, hid + 64 * i, valid);
        }
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * 128 + tid;
        const bool valid = i < n;
        float tcol[16];  // scaled dL/dh from the colour branch
#pragma unroll
        for (int j = 0; j < 16; j++) tcol[j] = 0.0f;
        float pre_ds = 0.0f, pre_h0 = 0.0f;
        if (with_rgb) {
            // ---- colour output layer: g3 (16) -> Ga (RB32), X = hid2
            {
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 8; j++) o[j] = 0;
                if (valid && dL_drgbs) {
                    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 3; j++) g[j] = pre_g[j] * (rgb_act ? pre_y[j] * (1.0f - pre_y[j]) : 1.0f) * loss_scale;
                    o[0] = pack2(g[0], g[1]); o[1] = pack2(g[2], 0.0f);
                }
                *reinterpret_cast<uint4*>(pGa + swz<32>(tid, 0)) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(pGa + swz<32>(tid, 1)) = make_uint4(o[4], o[5], o[6], o[7]);
                put_row128(xr);
            }
            layer(dGa32, 32, dX128, 128, wC3, 128, 64, 16, kColC3);
            fetch_row128(xr, hid1 + 64 * i, valid);
            if (valid && dL_dsigmas) { pre_ds = dL_dsigmas[i]; pre_h0 = h[16 * i]; }
            layer_wait();
            epilogue_mask64(trow + kColR, pX, pGb, tid);              // g2 -> Gb
            fence_before_sync(); __syncthreads();                     // everyone has read its X row before it is replaced
            put_row128(xr);
            layer(dGb128, 128, dX128, 128, wC2, 128, 64, 64, kColC2);
            fetch_row64(xs, in32 + 32 * i, valid);
            layer_wait();
            epilogue_mask64(trow + kColR, pX, pGa, tid);              // g1 -> Ga
            fence_before_sync(); __syncthreads();
            put_row64(xs);
            layer(dGa128, 128, dX64, 64, wC1, 64, 32, 64, kColC1);
            fetch_row128(xr, hid + 64 * i, valid);
            layer_wait();
            {   // R[:, 16:32] = scaled dL/dh from the colour branch
                tmem_ld16(trow + kColR + 16, tcol); tmem_ld_wait();
            }
        } else if (valid && dL_dsigmas) { pre_ds = dL_dsigmas[i]; pre_h0 = h[16 * i]; }
        // ---- density output layer: gh (16) -> Gb (RB32), X = hid
        {
            float g0 = tcol[0];
            if (valid && dL_dsigmas) g0 += (pre_ds * expf(fminf(fmaxf(pre_h0, -15.0f), 15.0f))) * loss_scale;
            uint32_t o[8];
            o[0] = pack2(g0, tcol[1]);
#pragma unroll
            for (int j = 1; j < 8; j++) o[j] = pack2(tcol[2 * j], tcol[2 * j + 1]);
            if (!valid) {
#pragma unroll
                for (int j = 0; j < 8; j++) o[j] = 0;
            }
            fence_before_sync(); __syncthreads();                     // previous layer's tiles are free
            *reinterpret_cast<uint4*>(pGb + swz<32>(tid, 0)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(pGb + swz<32>(tid, 1)) = make_uint4(o[4], o[5], o[6], o[7]);
            put_row128(xr);
        }
        layer(dGb32, 32, dX128, 128, wD2, 128, 64, 16, kColD2);
        fetch_row64(xs, feat + 32 * i, valid);
        {   // first-layer inputs of this CTA's next tile (xr is free: its row went to shared memory before this layer)
            const int64_t in = (tile + gridDim.x) * 128 + tid;
            const bool vn = tile + gridDim.x < n_tiles && in < n;
            if (with_rgb) {
                fetch_row128(xr, hid2 + 64 * in, vn);
#pragma unroll
                for (int j = 0; j < 3; j++) { pre_y[j] = vn ? rgbs[3 * in + j] : 0.0f; pre_g[j] = (vn && dL_drgbs) ? dL_drgbs[3 * in + j] : 0.0f; }
            } else {
                fetch_row128(xr, hid + 64 * in, vn);
            }
        }
        layer_wait();
        epilogue_mask64(trow + kColR, pX, pGa, tid);                  // gd -> Ga
        fence_before_sync(); __syncthreads();
        put_row64(xs);
        layer(dGa128, 128, dX64, 64, wD1, 64, 32, 64, kColD1);
        layer_wait();
        {
            float v[32];
            tmem_ld16(trow + kColR, v); tmem_ld16(trow + kColR + 16, v + 16); tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int q = 0; q < 8; q++)
                    reinterpret_cast<float4*>(dfeat + 32 * i)[q] =
                        make_float4(v[4 * q] * inv_scale, v[4 * q + 1] * inv_scale, v[4 * q + 2] * inv_scale, v[4 * q + 3] * inv_scale);
            }
        }
        acc = 1;
    }

    // ---- flush the weight gradients: M=64 accumulators live in lanes 0-15 of each warp's quadrant (row = 16*warp + lane).
    // Every CTA writes its five accumulators as one 10240-float slab of `wpart` (parameter order: colour 7168 | density
    // 3072) with plain stores; wgrad_reduce_kernel then sums the slabs in a fixed order.  (296 CTAs reducing onto the same
    // 10240 addresses with atomics serialised in L2 and cost 28 % of this kernel; the slab sum is also deterministic.)
    fence_before_sync(); __syncthreads(); fence_after_sync();
    {
        float* slab = wpart + (size_t)blockIdx.x * kWgradFloats;
        const int m = 16 * warp + lane;
        const bool own = lane < 16;
        auto flush = [&](uint32_t col, int ncols, float* dst, int ld_row, int ld_col, bool live) {
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                float v[16];
                if (live) { tmem_ld16(trow + col + c0, v); tmem_ld_wait(); }
                else {
#pragma unroll
                    for (int j = 0; j < 16; j++) v[j] = 0.0f;
                }
                if (own) {
                    if (ld_col == 1) {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            *reinterpret_cast<float4*>(dst + m * ld_row + c0 + 4 * q) =
                                make_float4(v[4 * q] * inv_scale, v[4 * q + 1] * inv_scale, v[4 * q + 2] * inv_scale, v[4 * q + 3] * inv_scale);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j++) dst[m * ld_row + (c0 + j) * ld_col] = v[j] * inv_scale;
                    }
                }
            }
        };
        const bool live_c = acc && with_rgb, live_d = acc != 0;
        flush(kColC3, 16, slab + 6144, 1, 64, live_c);   // accumulator is [in][out]: dW3[out][in] = acc[in][out]
        flush(kColC2, 64, slab + 2048, 64, 1, live_c);
        flush(kColC1, 32, slab, 32, 1, live_c);
        flush(kColD2, 16, slab + 7168 + 2048, 1, 64, live_d);
        flush(kColD1, 32, slab + 7168, 32, 1, live_d);
    }
    fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// grad += sum over the CTAs' slabs in a fixed order (deterministic).  A block owns 32 consecutive weights: warp w adds
// the slabs k = w, w+8, ... (128-byte coalesced rows), the eight partial sums are combined in warp order.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ wpart, int n_slabs, int with_rgb,
                                                           float* __restrict__ dWd, float* __restrict__ dWc) {
    __shared__ float part[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = blockIdx.x * 32 + lane;
    float s0 = 0.f, s1 = 0.f;
    int k = w;
    for (; k + 8 < n_slabs; k += 16) { s0 += wpart[(size_t)k * kWgradFloats + e]; s1 += wpart[(size_t)(k + 8) * kWgradFloats + e]; }
    if (k < n_slabs) s0 += wpart[(size_t)k * kWgradFloats + e];
    part[w][lane] = s0 + s1;
    __syncthreads();
    if (w == 0) {
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; j++) sum += part[j][lane];
        if (e < 7168) { if (with_rgb) dWc[e] += sum; } else dWd[e - 7168] += sum;
    }
}

}  // namespace arn

extern "C" ARN_API int arn_field_bw_tc(const float* xyzs, int64_t n, const float* xyz_min_host, const float* xyz_max_host, arn_levels_t levels,
                                       const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws,
                                       const float* sigmas, const float* rgbs, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale,
                                       float* dfeat_scratch, float* grad_params_xyz, float* grad_params_rgb, float* dL_dxyzs, arn_stream_t stream) {
    return arn_field_bw_tc_dyn(xyzs, n, nullptr, xyz_min_host, xyz_max_host, levels, params_xyz_f16, params_rgb_f16, rgb_act, ws, sigmas, rgbs, dL_dsigmas,
                               dL_drgbs, loss_scale, dfeat_scratch, grad_params_xyz, grad_params_rgb, dL_dxyzs, stream);
}
extern "C" ARN_API int arn_field_bw_tc_dyn(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                                           arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws,
                                           const float* sigmas, const float* rgbs, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale,
                                           float* dfeat_scratch, float* grad_params_xyz, float* grad_params_rgb, float* dL_dxyzs, arn_stream_t stream) {
    return arn::field_bw_tc_impl(xyzs, n, n_dev, xyz_min_host, xyz_max_host, levels, params_xyz_f16, params_rgb_f16, rgb_act, ws, sigmas, rgbs, dL_dsigmas,
                                 dL_drgbs, loss_scale, dfeat_scratch, grad_params_xyz, grad_params_rgb, dL_dxyzs, true, stream);
}
// pack_weights = false: ws.wimg still holds the weight image the forward of the same parameters built (fused training step).
int arn::field_bw_tc_impl(const float* xyzs, int64_t n, const int32_t* n_dev, const float* xyz_min_host, const float* xyz_max_host,
                          arn_levels_t levels, const void* params_xyz_f16, const void* params_rgb_f16, int rgb_act, arn_field_ws_t ws,
                          const float* sigmas, const float* rgbs, const float* dL_dsigmas, const float* dL_drgbs, float loss_scale,
                          float* dfeat_scratch, float* grad_params_xyz, float* grad_params_rgb, float* dL_dxyzs, bool pack_weights, arn_stream_t stream) {
    (void)sigmas;
    ARN_REQUIRE(n >= 0 && loss_scale > 0, "bad size / loss_scale");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(xyzs && params_xyz_f16 && ws.feat && ws.hid && ws.h && ws.wimg && dfeat_scratch && grad_params_xyz, "null pointer");
    const bool with_rgb = params_rgb_f16 != nullptr && dL_drgbs != nullptr;
    if (with_rgb) ARN_REQUIRE(ws.in32 && ws.hid1 && ws.hid2 && rgbs && grad_params_rgb, "null pointer (colour branch)");
    cudaStream_t st = (cudaStream_t)stream;
    const __half* pxyz = (const __half*)params_xyz_f16;
    // the weight image may have been built by a forward with different parameters only if the caller changed them in
    // between; rebuilding it here keeps the call self-contained (5 tiny blocks)
    if (pack_weights) {
        ARN_LAUNCH("pack_mlp_weights_kernel", st, arn::pack_mlp_weights_kernel<<<5, 256, 0, st>>>(pxyz, with_rgb ? (const __half*)params_rgb_f16 : nullptr, (uint8_t*)ws.wimg));
        if (int e = check_launch("pack_mlp_weights")) return e;
    }
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0; ARN_CUDA(cudaGetDevice(&dev)); ARN_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        ARN_CUDA(cudaFuncSetAttribute(arn::field_mlp_bw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, arn::kBwSmemBytes));
    }
    const int64_t n_tiles = (n + 127) / 128;
    int grid = (int)(n_tiles < (int64_t)n_sm * 2 ? n_tiles : (int64_t)n_sm * 2);
    if (grid > arn::kMaxBwCtas) grid = arn::kMaxBwCtas;
    float* wpart = reinterpret_cast<float*>((uint8_t*)ws.wimg + arn::kWimgBytes);  // slabs follow the weight image in the scratch
    ARN_LAUNCH("field_mlp_bw_tc_kernel", st, arn::field_mlp_bw_tc_kernel<<<grid, 128, arn::kBwSmemBytes, st>>>(
        n, n_dev, dL_dsigmas, with_rgb ? dL_drgbs : nullptr, rgbs, ws.h, (const __half*)ws.feat, (const __half*)ws.hid, (const __half*)ws.in32,
        (const __half*)ws.hid1, (const __half*)ws.hid2, (const uint8_t*)ws.wimg, rgb_act, with_rgb ? 1 : 0, loss_scale, grad_params_xyz,
        grad_params_rgb, dfeat_scratch, wpart));
    if (int e = check_launch("field_mlp_bw_tc")) return e;
    ARN_LAUNCH("wgrad_reduce_kernel", st, arn::wgrad_reduce_kernel<<<arn::kWgradFloats / 32, 256, 0, st>>>(wpart, grid, with_rgb ? 1 : 0, grad_params_xyz, grad_params_rgb));
    if (int e = check_launch("wgrad_reduce")) return e;
    return arn_hash_encode_bw_dyn(xyzs, n, n_dev, xyz_min_host, xyz_max_host, levels, pxyz + ARN_DENSITY_MLP_PARAMS, dfeat_scratch,
                                  grad_params_xyz + ARN_DENSITY_MLP_PARAMS, dL_dxyzs, stream);
}
