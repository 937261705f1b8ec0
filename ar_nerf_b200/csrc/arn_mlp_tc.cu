// libarnerf.so -- tensor-core MLPs of the field (density 32-64-16, colour 32-64-64-16) on tcgen05 / TMEM, fed by TMA.
//
// Activation images.  Every saved activation (feat, hid, in32, hid1, hid2: fp16; dfeat: fp32) lives in global memory as
// a sequence of 128-row TILES whose bytes are exactly the shared-memory operand image tcgen05 wants (rows of 64 or 128
// bytes, 16-byte chunks permuted by the hardware swizzle of that width: arn_field.cuh img_chunk64/128).  A tile is
// therefore ONE contiguous 8 or 16 KB block: the forward stores it and the backward loads it with a single bulk
// (TMA) copy issued by one thread -- no per-thread row loads/stores (which cost 32 L1 transactions per warp
// instruction and made both kernels LSU-bound), no tensor maps.  Buffers hold a whole number of tiles.
//
// Forward, one CTA of 128 threads per 128-sample tile (persistent over tiles, 2 CTAs per SM):
//   * the five weight matrices arrive once per CTA as ONE bulk copy of a pre-swizzled 20 KB operand image;
//   * the feature tile of the NEXT tile is requested (bulk copy + mbarrier) while the current one is processed;
//   * thread t owns sample row t: one elected thread issues the layer's tcgen05.mma chain (M=128, N=64|16, K=16 per
//     instruction, fp32 accumulators in TMEM) and commits to an mbarrier, every thread then pulls ITS row of the
//     accumulator with tcgen05.ld (32x32b: lane == row), applies the activation, rounds to fp16 and writes the next
//     layer's A tile; the same shared-memory tile is handed to the TMA engine as the saved activation.
//   Numeric contract: fp16 operands, fp32 accumulate -- identical rounding points to the simt
//   kernels and the oracle; only the accumulation order inside the MMA differs.
#include <mutex>

#include "arn_common.cuh"
#include "arn_field.cuh"
#include "arn_tc.cuh"

namespace arn {

using namespace tc;

// fp16 params -> swizzled operand image (arn_tc.cuh), stand-alone launch (the fused paths pack inside the hash-grid forward)
__global__ void __launch_bounds__(256) pack_mlp_weights_kernel(const __half* __restrict__ Wd, const __half* __restrict__ Wc,
                                                               uint8_t* __restrict__ img) {
    pack_weight_chunk(blockIdx.x * blockDim.x + threadIdx.x, Wd, Wc, img);
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

// number of pipelined tile ranges of a field evaluation over (up to) n samples: "pipeline_parts" (arn_set_tunable; default 1:
// on B200 the overlap of the L2-bound hash kernels with the MLP kernels loses to the fixed cost every extra persistent
// launch pays -- TMEM allocation, weight image, the 40 KB weight-gradient slab per CTA); 1 for small calls in any case
inline int pipeline_parts(int64_t n) {
    const int want = tunable(kTunPipelineParts);
    if (want <= 1 || n < 64 * 1024) return 1;
    return want > 8 ? 8 : want;
}

constexpr int kFwSmemTile32 = 128 * 64;    // 128 rows x 32 halves
constexpr int kFwSmemTile64 = 128 * 128;   // 128 rows x 64 halves
// Shared-memory map of the forward (offsets from the 1024-aligned base): weight image | per warpgroup: F (the tile's
// features) | A B (two 16 KB activation buffers used in turn: hid -> A, in32 -> B, hid1 -> A, hid2 -> B).  The h staging
// tile (128 x 16 f32, 8 KB) sits in the upper half of B beside in32 (8 KB): both are gone when hid2 arrives.  F is free as
// soon as the first MMA of a tile has read it: the NEXT tile's features are requested then and have four layers to arrive.
// A buffer is overwritten two layers after it was written; its bulk store (issued one layer after the write) has then had
// a whole layer to read it, and the thread that issues the MMAs waits for that read before it issues the layer whose
// epilogue overwrites the buffer.
// The kernel is a chain of short latencies per tile (MMA completion, TMEM load, shared-memory fence: issue slots 30 % busy,
// tensor pipe 10 %), so what counts is the number of tiles in flight on an SM and the number of ROUNDS a launch takes:
// W warpgroups per CTA share ONE weight image, every warpgroup owns the three tile buffers above (40 KB) and 64 TMEM
// columns (a layer's accumulator has been pulled before the next layer is issued) and walks its own tiles.
// W = 5 (640 threads, one CTA per SM: 20 + 5 x 40 KB): a training step's 1920 tiles take 3 rounds over 740 warpgroups where
// four warpgroups of 48 KB took 4 over 592; W = 1 (three CTAs per SM) for launches too small to fill 5 x n_sm tiles.
constexpr int kFwF = 0, kFwA = kFwF + kFwSmemTile32, kFwB = kFwA + kFwSmemTile64;  // offsets inside a warpgroup's block
constexpr int kFwWgBytes = kFwB + kFwSmemTile64;
constexpr int kFwWide = 5;
template <int W> constexpr int fw_smem_bytes() { return kWimgBytes + W * kFwWgBytes + 1024; }  // + alignment slack
template <int W> constexpr int fw_tmem_cols() { return W == 1 ? 64 : 512; }                     // 64 per warpgroup, a power of two
constexpr int kFwCtasPerSm = 3;  // W = 1

template <int W>
__global__ void __launch_bounds__(128 * W, 1) field_mlp_fw_tc_kernel(const __half* __restrict__ feat, const float* __restrict__ dirs, int64_t n,
                                                              const int32_t* __restrict__ n_dev, const uint8_t* __restrict__ wimg, int rgb_act, int with_rgb,
                                                              __half* __restrict__ hid, float* __restrict__ h, float* __restrict__ sigmas,
                                                              __half* __restrict__ in32, __half* __restrict__ hid1, __half* __restrict__ hid2,
                                                              float* __restrict__ rgbs, int part, int parts) {
    pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[1 + 2 * W];  // weights | per warpgroup: mma, features
    __shared__ uint32_t tmem_slot;
    if (n_dev) n = min(n, (int64_t)*n_dev);  // fused step: the sample count lives on the device
    const uint32_t base0 = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzled tiles need 1024-byte alignment
    const int wg = threadIdx.x >> 7, tid = threadIdx.x & 127, warp = tid >> 5;
    const uint32_t sW = base0;
    const uint32_t base = base0 + kWimgBytes + wg * kFwWgBytes;  // this warpgroup's F A B
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_w = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1 + 2 * wg]), bar_f = smem_u32(&bars[2 + 2 * wg]);

    if (threadIdx.x == 0) {
        for (int k = 0; k < 1 + 2 * W; k++) mbar_init(smem_u32(&bars[k]), 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, fw_tmem_cols<W>());
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tmem = tmem_slot + 64 * wg;  // this warpgroup's 64 columns
    const uint32_t wbytes = with_rgb ? kWimgBytes : kWimgC1;
    // this launch owns the tiles [t_begin, n_tiles) of the sample list (one of `parts` consecutive ranges: the host pipelines
    // the hash-grid forward of range p+1 on a second stream under the MLP of range p)
    const int64_t n_tiles_all = (n + 127) / 128;
    const int64_t t_begin = n_tiles_all * part / parts, n_tiles = n_tiles_all * (part + 1) / parts;
    const int64_t tile0 = t_begin + (int64_t)blockIdx.x * W + wg, tile_stride = (int64_t)gridDim.x * W;
    if (threadIdx.x == 0) { mbar_expect_tx(bar_w, wbytes); bulk_g2s(sW, wimg, wbytes, bar_w); }
    if (tid == 0 && tile0 < n_tiles) { mbar_expect_tx(bar_f, kFwSmemTile32); bulk_g2s(base + kFwF, feat + tile0 * 128 * 32, kFwSmemTile32, bar_f); }
    mbar_wait(bar_w, 0);

    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes; lane == sample row
    const uint64_t aF = smem_desc<64>(base + kFwF);
    const uint64_t aI = smem_desc<64>(base + kFwB);
    const uint64_t aH0 = smem_desc<128>(base + kFwA), aH1 = smem_desc<128>(base + kFwA), aH2 = smem_desc<128>(base + kFwB);
    constexpr int kFwH0 = kFwA, kFwI = kFwB, kFwH1 = kFwA, kFwH2 = kFwB;
    constexpr int kFwHS = kFwB + kFwSmemTile32;  // h staging: upper half of B, beside in32
    const uint64_t bD1 = smem_desc<64>(sW + kWimgD1), bD2 = smem_desc<128>(sW + kWimgD2);
    const uint64_t bC1 = smem_desc<64>(sW + kWimgC1), bC2 = smem_desc<128>(sW + kWimgC2), bC3 = smem_desc<128>(sW + kWimgC3);
    constexpr uint32_t kI64 = instr_desc(128, 64, 0, 0), kI16 = instr_desc(128, 16, 0, 0);
    uint32_t phase = 0;

    // issue one layer: D[tmem cols] = A (128 x K) * W^T, K = 16 * ksteps; then commit.  The barrier in front publishes the
    // A tile the threads have just written (generic proxy -> async proxy) to the tensor core AND to the TMA engine, and
    // orders the threads' tcgen05.ld of the previous accumulator before the MMAs that overwrite it.
    // free_buf: the epilogue of this layer overwrites a buffer that an earlier bulk store may still be reading; the issuing
    // thread waits for those reads BEFORE it issues the MMAs, and the other threads cannot write before the MMAs commit.
    auto issue_only = [&](uint64_t a, uint64_t b, uint32_t idesc, int ksteps, bool free_buf) {
        fence_before_sync(); fence_async_smem(); wg_sync(1 + wg);
        if (tid == 0) {
            fence_after_sync();
            if (free_buf) bulk_wait_read<0>();
            for (int k = 0; k < ksteps; k++) mma_f16(tmem, desc_advance(a, 32 * k), desc_advance(b, 32 * k), idesc, k > 0);
            mma_commit(bar_mma);
        }
    };
    auto wait_mma = [&]() {
        mbar_wait(bar_mma, phase); phase ^= 1;
        fence_after_sync();
    };
    // epilogue of a 64-wide hidden layer: ReLU, fp16, this thread's row of the next A tile (two halves of 32 columns: the
    // first half's conversions and stores run under the second half's TMEM load)
    auto hidden_row = [&](uint8_t* tile) {
        float v[64];
#pragma unroll
        for (int q = 0; q < 4; q++) tmem_ld16(trow + 16 * q, v + 16 * q);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; c++) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; j++) o[j] = pack2(fmaxf(v[8 * c + 2 * j], 0.0f), fmaxf(v[8 * c + 2 * j + 1], 0.0f));
            *reinterpret_cast<uint4*>(tile + swz<128>(tid, c)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    };

    // ray direction of this thread's sample, requested one tile ahead (between an MMA issue and its completion)
    float dr[3] = {1.0f, 0.0f, 0.0f};
    auto fetch_dir = [&](int64_t t) {
        const int64_t i = t * 128 + tid;
        const bool ok = with_rgb && t < n_tiles && i < n;
        dr[0] = ok ? dirs[3 * i] : 1.0f; dr[1] = ok ? dirs[3 * i + 1] : 0.0f; dr[2] = ok ? dirs[3 * i + 2] : 0.0f;
    };
    fetch_dir(tile0);
    int it = 0;
    for (int64_t tile = tile0; tile < n_tiles; tile += tile_stride, it++) {
        const int64_t i = tile * 128 + tid;
        const bool valid = i < n;
        uint8_t* pF = sm + kFwF;
        mbar_wait(bar_f, (uint32_t)it & 1u);
        if (!valid) {  // rows past the sample count: zero features keep every saved activation of the pad rows finite
#pragma unroll
            for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(pF + swz<64>(tid, c)) = make_uint4(0, 0, 0, 0);
        }
        const float dcur[3] = {dr[0], dr[1], dr[2]};
        // ---- density layer 1: A = feature tile
        issue_only(aF, bD1, kI64, 2, true);                      // epilogue -> A (previous tile's hid1)
        wait_mma();
        if (tid == 0) {  // F has been read: the next tile's features may land in it
            const int64_t next = tile + tile_stride;
            if (next < n_tiles) { mbar_expect_tx(bar_f, kFwSmemTile32); bulk_g2s(base + kFwF, feat + next * 128 * 32, kFwSmemTile32, bar_f); }
        }
        hidden_row(sm + kFwH0);
        // ---- density layer 2 -> h (16, fp32), sigma
        issue_only(aH0, bD2, kI16, 4, true);                     // epilogue -> B (previous tile's hid2): in32 and the h staging tile
        if (tid == 0 && hid) { bulk_s2g(hid + tile * 128 * 64, base + kFwH0, kFwSmemTile64); bulk_commit(); }
        wait_mma();
        {
            float hv[16];
            tmem_ld16(trow, hv); tmem_ld_wait();
            if (valid) sigmas[i] = expf(hv[0]);
            // h tile staging: plain row-major 128 x 16 f32, the public layout of h
            if (h) {
#pragma unroll
                for (int c = 0; c < 4; c++)
                    *reinterpret_cast<float4*>(sm + kFwHS + tid * 64 + c * 16) = make_float4(hv[4 * c], hv[4 * c + 1], hv[4 * c + 2], hv[4 * c + 3]);
            }
            if (with_rgb) {  // colour-net input row [sh16 | fp16(h16)]
                float sh[16];
                sh4_eval(dcur, sh);
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 8; j++) { o[j] = pack2(sh[2 * j], sh[2 * j + 1]); o[8 + j] = pack2(hv[2 * j], hv[2 * j + 1]); }
#pragma unroll
                for (int c = 0; c < 4; c++) *reinterpret_cast<uint4*>(sm + kFwI + swz<64>(tid, c)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
            }
        }
        if (!with_rgb) {
            if (h) {
                fence_before_sync(); fence_async_smem(); wg_sync(1 + wg);
                if (tid == 0) { bulk_s2g(h + tile * 128 * 16, base + kFwHS, 128 * 64); bulk_commit(); }
            }
            continue;
        }
        // ---- colour layer 1
        issue_only(aI, bC1, kI64, 2, true);                      // epilogue -> A (hid)
        if (tid == 0) {
            if (hid) bulk_s2g(in32 + tile * 128 * 32, base + kFwI, kFwSmemTile32);
            if (h) bulk_s2g(h + tile * 128 * 16, base + kFwHS, 128 * 64);
            bulk_commit();
        }
        wait_mma();
        hidden_row(sm + kFwH1);
        // ---- colour layer 2
        issue_only(aH1, bC2, kI64, 4, true);                     // epilogue -> B (in32, h staging)
        if (tid == 0 && hid) { bulk_s2g(hid1 + tile * 128 * 64, base + kFwH1, kFwSmemTile64); bulk_commit(); }
        wait_mma();
        hidden_row(sm + kFwH2);
        // ---- colour layer 3 -> rgb
        issue_only(aH2, bC3, kI16, 4, false);
        if (tid == 0 && hid) { bulk_s2g(hid2 + tile * 128 * 64, base + kFwH2, kFwSmemTile64); bulk_commit(); }
        fetch_dir(tile + tile_stride);
        wait_mma();
        {
            float ov[16];
            tmem_ld16(trow, ov); tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int j = 0; j < 3; j++) rgbs[3 * i + j] = rgb_act ? 1.0f / (1.0f + expf(-ov[j])) : ov[j];
            }
        }
    }
    if (tid == 0) bulk_wait_read<0>();
    fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_slot, fw_tmem_cols<W>());
}

// One-time, per device: dynamic shared memory limits of every instance of the two kernels.  Returns the SM count.
int mlp_device_setup(int* n_sm_out);

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
    ARN_REQUIRE(xyzs && params_xyz_f16 && ws.feat && ws.wimg && sigmas, "null pointer");
    const bool with_rgb = dirs != nullptr;
    // ws.hid == NULL: inference -- no activation is saved for a backward (ws.h is optional in either mode)
    if (with_rgb) ARN_REQUIRE(params_rgb_f16 && rgbs && (!ws.hid || (ws.in32 && ws.hid1 && ws.hid2)), "null pointer (colour branch)");
    cudaStream_t st = (cudaStream_t)stream;
    const __half* pxyz = (const __half*)params_xyz_f16;
    int n_sm = 0;
    if (int e = arn::mlp_device_setup(&n_sm)) return e;
    // Pipelined over `parts` consecutive ranges of tiles: the hash-grid forward (gathers out of L2) of range p+1 runs on
    // a second stream under the MLP (tensor core + HBM stores) of range p -- the two kernels stress different units.
    const int64_t n_tiles = (n + 127) / 128;
    const int parts = pipeline_parts(n);
    // kFwWide warpgroups per CTA, one CTA per SM, when the launch has the tiles to fill that; else one warpgroup per CTA
    const bool wide = (n_tiles + parts - 1) / parts >= (int64_t)n_sm * kFwWide && (tunable(kTunMlpWide) & 1) != 0;
    const int wgs = wide ? kFwWide : 1;
    const int grid = (int)max((int64_t)1, min((int64_t)n_sm * (wide ? 1 : kFwCtasPerSm), ((n_tiles + parts - 1) / parts + wgs - 1) / wgs));
    PipeStreams* ps = nullptr;
    if (parts > 1) {
        if (int e = pipe_streams(&ps)) return e;
        ARN_CUDA(cudaEventRecord(ps->fork, st));
        ARN_CUDA(cudaStreamWaitEvent(ps->side, ps->fork, 0));
    }
    for (int p = 0; p < parts; p++) {
        cudaStream_t hs = parts > 1 ? ps->side : st;
        // the weight image is packed by the first blocks of the first hash-grid launch (no launch of its own)
        if (int e = hash_encode_fw_impl(xyzs, n, n_dev, xyz_min_host, xyz_max_host, levels, pxyz + ARN_DENSITY_MLP_PARAMS, ws.feat, /*tile_image=*/1, hs,
                                        p == 0 ? pxyz : nullptr, p == 0 ? (const __half*)params_rgb_f16 : nullptr, p == 0 ? (uint8_t*)ws.wimg : nullptr,
                                        p, parts)) return e;
        if (parts > 1) ARN_CUDA(cudaEventRecord(ps->ev[p], hs));
    }
    for (int p = 0; p < parts; p++) {
        if (parts > 1) ARN_CUDA(cudaStreamWaitEvent(st, ps->ev[p], 0));
#define ARN_FW_ARGS (const __half*)ws.feat, dirs, n, n_dev, (const uint8_t*)ws.wimg, rgb_act, with_rgb ? 1 : 0, (__half*)ws.hid, ws.h, sigmas, \
            (__half*)ws.in32, (__half*)ws.hid1, (__half*)ws.hid2, rgbs, p, parts
        if (wide) ARN_LAUNCH_PDL("field_mlp_fw_tc_kernel", st, (field_mlp_fw_tc_kernel<kFwWide>), grid, 128 * kFwWide, fw_smem_bytes<kFwWide>(), ARN_FW_ARGS);
        else ARN_LAUNCH_PDL("field_mlp_fw_tc_kernel", st, (field_mlp_fw_tc_kernel<1>), grid, 128, fw_smem_bytes<1>(), ARN_FW_ARGS);
#undef ARN_FW_ARGS
        if (int e = check_launch("field_mlp_fw_tc")) return e;
    }
    return ARN_OK;
}

namespace arn {
using namespace tc;
// =====================================================================================================================
// Backward.  W warpgroups of 128 threads per CTA, each walking its own sequence of 128-sample tiles (thread t of a warpgroup
// = sample row t), persistent; ONE CTA per SM (W = 3: 384 threads) or, for launches too small to fill that, W = 1.
// Per layer, walking the net backwards, ONE commit covers two MMA chains that read the same two shared-memory tiles:
//   dgrad   R[128 x in]  = G (K-major A: rows = samples, K = out)  x  W (MN-major B straight from the forward's weight image)
//   wgrad   dW[out x in] += G^T X : A = G as MN-major (M = out), B = X as MN-major (N = in), K = the 128 samples of the tile
// The five weight-gradient accumulators stay in TENSOR MEMORY for the whole kernel (160 columns, zero-initialised with
// tcgen05.st, SHARED by the CTA's warpgroups: the tensor pipe executes their MMAs in issue order) and are flushed once per
// CTA; layers with 16 outputs accumulate the transposed product (M = in = 64, N = 16).
// Thread t then pulls row t of R with tcgen05.ld, applies the ReLU mask of its own activation row, scales/rounds to
// fp16 and writes the next G tile.  Rounding points are those of the simt kernel / oracle (fp16 G, fp32 accumulate).
// The saved activation tiles X (hid2, hid1, in32, hid, feat -- in this order, 5 per sample tile) stream through a ring of
// three 16 KB slots per warpgroup: one thread issues each tile as a single bulk (TMA) copy two layers before its MMA, as
// soon as the slot's previous tile has been consumed; the layer's threads only wait on the slot's mbarrier.  G lives in ONE
// 16 KB buffer (the MMA that read G_k has completed before the epilogue writes G_k+1); dfeat is staged in the ring slot the
// tile's last X (feat) has just left and stored by TMA.  Shared memory: weight image 20 KB + W x 64 KB; the latency chain
// of a tile (issue -> commit -> tcgen05.ld -> mask -> store, five times) is hidden by the W tiles in flight on the SM --
// three instead of the two that two 100 KB CTAs gave (the kernel is latency-bound: tensor pipe 13 %, DRAM 31 %).
// TMEM map: [64 wg, 64 wg + 64) R of warpgroup wg | then dWc3^T (16) dWc2 (64) dWc1 (32) dWd2^T (16) dWd1 (32)
constexpr int kBwWgBytes = 4 * kFwSmemTile64;  // per warpgroup: G | X0 X1 X2
template <int W> constexpr int bw_smem_bytes() { return kWimgBytes + W * kBwWgBytes + 1024; }
constexpr int kMaxBwCtas = ARN_FIELD_SCRATCH_SLABS;

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

template <int W>
__global__ void __launch_bounds__(128 * W, 1) __maxnreg__(W == 3 ? 128 : 192) field_mlp_bw_tc_kernel(int64_t n, const int32_t* __restrict__ n_dev, const float* __restrict__ dL_dsigmas, const float* __restrict__ dL_drgbs,
                                                              const float* __restrict__ rgbs, const float* __restrict__ sigmas,
                                                              const __half* __restrict__ feat, const __half* __restrict__ hid,
                                                              const __half* __restrict__ in32, const __half* __restrict__ hid1,
                                                              const __half* __restrict__ hid2, const uint8_t* __restrict__ wimg, int rgb_act,
                                                              int with_rgb, float loss_scale, float exp_hi, float* __restrict__ dfeat, float* __restrict__ wpart,
                                                              int part, int parts, int slab0, int reverse) {
    pdl_enter();
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[1 + 4 * W];  // weights | per warpgroup: mma, x0, x1, x2
    __shared__ uint32_t tmem_slot;
    constexpr uint32_t kTmemCols = W == 1 ? 256 : 512;
    constexpr uint32_t kColW = 64 * W, kColC3 = kColW, kColC2 = kColW + 16, kColC1 = kColW + 80, kColD2 = kColW + 112, kColD1 = kColW + 128;
    if (n_dev) n = min(n, (int64_t)*n_dev);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int wg = threadIdx.x >> 7, tid = threadIdx.x & 127, warp = tid >> 5, lane = tid & 31;
    const uint32_t sW = base;
    const uint32_t sWg = base + kWimgBytes + wg * kBwWgBytes;      // this warpgroup's G | X0 X1 X2
    uint8_t* pWg = sm + kWimgBytes + wg * kBwWgBytes;
    const uint32_t sG = sWg; uint8_t* pG = pWg;
    constexpr int kX0 = kFwSmemTile64;
    const uint32_t bar_w = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1 + 4 * wg]);
    auto bar_x = [&](int slot) { return smem_u32(&bars[2 + 4 * wg + slot]); };

    if (threadIdx.x == 0) {
        for (int k = 0; k < 1 + 4 * W; k++) mbar_init(smem_u32(&bars[k]), 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&tmem_slot, kTmemCols);
    fence_before_sync(); __syncthreads(); fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's 32 TMEM lanes; lane == sample row of the warpgroup's tile
    // zero the weight-gradient accumulators (every warpgroup's MMAs accumulate into them from its first tile on)
    if (wg == 0) {
#pragma unroll
        for (int c = 0; c < 160; c += 16) tmem_st16_fill(trow + kColW + c, 0u);
        tmem_st_wait();
    }
    const uint32_t wbytes = with_rgb ? kWimgBytes : kWimgC1;
    // tiles [t_begin, n_tiles) of the sample list (see the forward kernel); warpgroup wg of CTA b takes t_begin + b W + wg, ...
    const int64_t n_tiles_all = (n + 127) / 128;
    const int64_t t_begin = n_tiles_all * part / parts, n_tiles = n_tiles_all * (part + 1) / parts;
    const int64_t tile0 = t_begin + (int64_t)blockIdx.x * W + wg, tile_stride = (int64_t)gridDim.x * W;
    // reverse: position t of the walk is the tile n_tiles - 1 - (t - t_begin) -- the forward's LAST tiles first, whose saved
    // activations are the ones still in L2 (the forward wrote 110 MB through a 126 MB cache in tile order)
    auto phys = [&](int64_t t) { return reverse ? n_tiles - 1 - (t - t_begin) : t; };

    // ---- activation-tile stream: load j of this warpgroup = tile (j / L) of its tile sequence, kind (j % L); slot j % 3
    const int L = with_rgb ? 5 : 2;
    auto issue_load = [&](int64_t j) {  // the warpgroup's thread 0 only
        int64_t t = tile0 + (j / L) * tile_stride;
        if (t >= n_tiles) return;
        t = phys(t);
        const int kind = with_rgb ? (int)(j % 5) : 3 + (int)(j % 2);
        const __half* src; uint32_t bytes;
        switch (kind) {
            case 0: src = hid2 + t * 128 * 64; bytes = kFwSmemTile64; break;
            case 1: src = hid1 + t * 128 * 64; bytes = kFwSmemTile64; break;
            case 2: src = in32 + t * 128 * 32; bytes = kFwSmemTile32; break;
            case 3: src = hid + t * 128 * 64; bytes = kFwSmemTile64; break;
            default: src = feat + t * 128 * 32; bytes = kFwSmemTile32; break;
        }
        const int slot = (int)(j % 3);
        mbar_expect_tx(bar_x(slot), bytes);
        bulk_g2s(sWg + kX0 + slot * kFwSmemTile64, src, bytes, bar_x(slot));
    };
    if (threadIdx.x == 0) { mbar_expect_tx(bar_w, wbytes); bulk_g2s(sW, wimg, wbytes, bar_w); }
    if (tid == 0) { issue_load(0); issue_load(1); }
    fence_before_sync(); __syncthreads(); fence_after_sync();  // the zeroed accumulators are visible to every issuing thread
    mbar_wait(bar_w, 0);

    uint32_t phase = 0;
    const float inv_scale = 1.0f / loss_scale;
    const float exp_lo = 1.0f / exp_hi;
    int64_t j = 0;     // index of the next activation tile to be consumed

    // One layer step: wait for the layer's X tile, publish the G tile, then wgrad (K = 128 samples, 8 MMAs) + dgrad
    // (K = out, ksteps MMAs) under one commit.  After the barrier the slot of the PREVIOUS layer's X tile is free (its MMA
    // has completed and every thread is past its epilogue), so thread 0 refills it with the tile two layers ahead -- after
    // the dfeat store that may have been staged there has read it.
    //   g   : G tile (rows = samples, RBG bytes per row)        x : X tile in slot j % 3 (RBX bytes per row)
    //   w   : weight tile of this layer in the image (MN-major B for dgrad), RBW bytes per row (= 2*in)
    //   wgrad M=64: A = (t_out16 ? x : g) MN-major, B = (t_out16 ? g : x) MN-major
    auto layer = [&](uint64_t g_desc, int rbg, int rbx, uint64_t w_desc, int rbw, int n_in, int n_out, uint32_t wcol) -> const uint8_t* {
        const int slot = (int)(j % 3);
        const uint32_t sX = sWg + kX0 + slot * kFwSmemTile64;
        mbar_wait(bar_x(slot), (uint32_t)(j / 3) & 1u);
        fence_before_sync(); fence_async_smem(); wg_sync(1 + wg);
        if (tid == 0) {
            fence_after_sync();
            bulk_wait_read<0>();
            issue_load(j + 2);
            const uint64_t x_desc = rbx == 128 ? smem_desc<128>(sX) : smem_desc<64>(sX);
            const bool t16 = n_out == 16;
            const uint64_t wa = t16 ? x_desc : g_desc, wb = t16 ? g_desc : x_desc;
            const int rba = t16 ? rbx : rbg, rbb = t16 ? rbg : rbx;
            const uint32_t wi = instr_desc(64, t16 ? 16 : n_in, 1, 1);
            for (int k = 0; k < 8; k++)  // 16 samples per MMA = 16 rows of each tile
                mma_f16(tmem + wcol, desc_advance(wa, 16 * rba * k), desc_advance(wb, 16 * rbb * k), wi, 1u);
            const uint32_t di = instr_desc(128, n_in, 0, 1);
            for (int k = 0; k < n_out / 16; k++)  // K = out: 32 B along a G row, 16 rows down the weight tile
                mma_f16(tmem + 64 * wg, desc_advance(g_desc, 32 * k), desc_advance(w_desc, 16 * rbw * k), di, k > 0);
            mma_commit(bar_mma);
        }
        j++;
        return pWg + kX0 + slot * kFwSmemTile64;
    };
    auto layer_wait = [&]() {
        mbar_wait(bar_mma, phase); phase ^= 1;
        fence_after_sync();
    };
    const uint32_t tR = trow + 64 * wg;  // this thread's row of the warpgroup's R

    const uint64_t dG128 = smem_desc<128>(sG), dG32 = smem_desc<32>(sG);
    const uint64_t wD1 = smem_desc<64>(sW + kWimgD1), wD2 = smem_desc<128>(sW + kWimgD2);
    const uint64_t wC1 = smem_desc<64>(sW + kWimgC1), wC2 = smem_desc<128>(sW + kWimgC2), wC3 = smem_desc<128>(sW + kWimgC3);

    // per-sample scalars, requested one tile ahead between an MMA issue and its completion (ptxas hoists loads to the top
    // of a basic block but not across the bar.sync; a use placed after younger loads would wait for those as well)
    float pre_y[3] = {0.f, 0.f, 0.f}, pre_g[3] = {0.f, 0.f, 0.f}, pre_ds = 0.f, pre_sig = 1.f;
    auto fetch_scalars = [&](int64_t t) {
        const int64_t i = phys(t) * 128 + tid;
        const bool ok = t < n_tiles && i < n;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            pre_y[k] = (ok && with_rgb) ? rgbs[3 * i + k] : 0.0f;
            pre_g[k] = (ok && with_rgb && dL_drgbs) ? dL_drgbs[3 * i + k] : 0.0f;
        }
        pre_ds = (ok && dL_dsigmas) ? dL_dsigmas[i] : 0.0f;
        pre_sig = ok ? sigmas[i] : 1.0f;
    };
    fetch_scalars(tile0);

    for (int64_t pos = tile0; pos < n_tiles; pos += tile_stride) {
        const int64_t tile = phys(pos);
        const int64_t i = tile * 128 + tid;
        const bool valid = i < n;
        float tcol[16];  // scaled dL/dh from the colour branch
#pragma unroll
        for (int k = 0; k < 16; k++) tcol[k] = 0.0f;
        // d sigma / d h0 = exp(clamp(h0, -15, 15)) (custom_functions.py:170-173) = clamp(sigma, e^-15, e^15): exp is monotone
        const float g_sigma = valid ? pre_ds * fminf(fmaxf(pre_sig, exp_lo), exp_hi) * loss_scale : 0.0f;
        // G was last read by the previous tile's D1 MMA, which has completed (layer_wait); the dfeat staging does not use it
        if (with_rgb) {
            // ---- colour output layer: g3 (16) -> G (RB32), X = hid2
            {
                uint32_t o[8];
#pragma unroll
                for (int k = 0; k < 8; k++) o[k] = 0;
                if (valid) {
                    float g[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) g[k] = pre_g[k] * (rgb_act ? pre_y[k] * (1.0f - pre_y[k]) : 1.0f) * loss_scale;
                    o[0] = pack2(g[0], g[1]); o[1] = pack2(g[2], 0.0f);
                }
                *reinterpret_cast<uint4*>(pG + swz<32>(tid, 0)) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(pG + swz<32>(tid, 1)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
            const uint8_t* x;
            x = layer(dG32, 32, 128, wC3, 128, 64, 16, kColC3);
            layer_wait();
            epilogue_mask64(tR, x, pG, tid);                           // g2 -> G
            x = layer(dG128, 128, 128, wC2, 128, 64, 64, kColC2);
            layer_wait();
            epilogue_mask64(tR, x, pG, tid);                           // g1 -> G
            layer(dG128, 128, 64, wC1, 64, 32, 64, kColC1);
            layer_wait();
            tmem_ld16(tR + 16, tcol); tmem_ld_wait();                   // R[:, 16:32] = scaled dL/dh from the colour branch
        }
        // ---- density output layer: gh (16) -> G (RB32), X = hid
        {
            uint32_t o[8];
            o[0] = pack2(tcol[0] + g_sigma, tcol[1]);
#pragma unroll
            for (int k = 1; k < 8; k++) o[k] = pack2(tcol[2 * k], tcol[2 * k + 1]);
            if (!valid) {
#pragma unroll
                for (int k = 0; k < 8; k++) o[k] = 0;
            }
            // G was last read by the C1 MMA (or the previous tile's D1 MMA), which has completed
            *reinterpret_cast<uint4*>(pG + swz<32>(tid, 0)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(pG + swz<32>(tid, 1)) = make_uint4(o[4], o[5], o[6], o[7]);
        }
        const uint8_t* xh = layer(dG32, 32, 128, wD2, 128, 64, 16, kColD2);
        layer_wait();
        epilogue_mask64(tR, xh, pG, tid);                              // gd -> G
        const uint8_t* xf = layer(dG128, 128, 64, wD1, 64, 32, 64, kColD1);
        fetch_scalars(pos + tile_stride);
        layer_wait();
        {   // dfeat tile: fp32 rows of 128 B, chunk-permuted like a RB128 image; staged in the ring slot the feature tile has
            // just left (free until the load two layers ahead, which waits for this store's read) and stored by TMA
            uint8_t* stage = const_cast<uint8_t*>(xf);
            float v[32];
            tmem_ld16(tR, v); tmem_ld16(tR + 16, v + 16); tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; q++)
                *reinterpret_cast<float4*>(stage + swz<128>(tid, q)) =
                    make_float4(v[4 * q] * inv_scale, v[4 * q + 1] * inv_scale, v[4 * q + 2] * inv_scale, v[4 * q + 3] * inv_scale);
            fence_async_smem(); wg_sync(1 + wg);
            if (tid == 0) { bulk_s2g(dfeat + tile * 128 * 32, sWg + kX0 + (uint32_t)(stage - (pWg + kX0)), kFwSmemTile64); bulk_commit(); }
        }
    }
    if (tid == 0) bulk_wait_read<0>();

    // ---- flush the weight gradients: M=64 accumulators live in lanes 0-15 of each warp's quadrant (row = 16*warp + lane).
    // Every CTA writes its five accumulators as one 10240-float slab of `wpart` (parameter order: colour 7168 | density
    // 3072) with plain stores; wgrad_reduce_kernel then sums the slabs in a fixed order.  (296 CTAs reducing onto the same
    // 10240 addresses with atomics serialised in L2 and cost 28 % of this kernel; the slab sum is also deterministic.)
    fence_before_sync(); __syncthreads(); fence_after_sync();
    if (wg == 0) {
        float* slab = wpart + (size_t)(slab0 + blockIdx.x) * kWgradFloats;
        const int m = 16 * warp + lane;
        const bool own = lane < 16;
        auto flush = [&](uint32_t col, int ncols, float* dst, int ld_row, int ld_col, bool live) {
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                float v[16];
                if (live) { tmem_ld16(trow + col + c0, v); tmem_ld_wait(); }
                else {
#pragma unroll
                    for (int jj = 0; jj < 16; jj++) v[jj] = 0.0f;
                }
                if (own) {
                    if (ld_col == 1) {
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            *reinterpret_cast<float4*>(dst + m * ld_row + c0 + 4 * q) =
                                make_float4(v[4 * q] * inv_scale, v[4 * q + 1] * inv_scale, v[4 * q + 2] * inv_scale, v[4 * q + 3] * inv_scale);
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 16; jj++) dst[m * ld_row + (c0 + jj) * ld_col] = v[jj] * inv_scale;
                    }
                }
            }
        };
        const bool live_c = with_rgb != 0;
        flush(kColC3, 16, slab + 6144, 1, 64, live_c);   // accumulator is [in][out]: dW3[out][in] = acc[in][out]
        flush(kColC2, 64, slab + 2048, 64, 1, live_c);
        flush(kColC1, 32, slab, 32, 1, live_c);
        flush(kColD2, 16, slab + 7168 + 2048, 1, 64, true);
        flush(kColD1, 32, slab + 7168, 32, 1, true);
    }
    fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, kTmemCols);
}

int mlp_device_setup(int* n_sm_out) {
    static int n_sm[16] = {};
    static std::mutex mu;
    int dev = 0;
    ARN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) { set_error("mlp_device_setup: device index out of range"); return ARN_E_INVALID; }
    std::lock_guard<std::mutex> lk(mu);
    if (!n_sm[dev]) {
        int n = 0;
        ARN_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        ARN_CUDA(cudaFuncSetAttribute(field_mlp_fw_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, fw_smem_bytes<1>()));
        ARN_CUDA(cudaFuncSetAttribute(field_mlp_fw_tc_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ARN_CUDA(cudaFuncSetAttribute(field_mlp_fw_tc_kernel<kFwWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, fw_smem_bytes<kFwWide>()));
        ARN_CUDA(cudaFuncSetAttribute(field_mlp_bw_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw_smem_bytes<1>()));
        ARN_CUDA(cudaFuncSetAttribute(field_mlp_bw_tc_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ARN_CUDA(cudaFuncSetAttribute(field_mlp_bw_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw_smem_bytes<3>()));
        n_sm[dev] = n;
    }
    *n_sm_out = n_sm[dev];
    return ARN_OK;
}

// grad += sum over the CTAs' slabs (arn_tc.cuh wgrad_reduce_block), stand-alone launch
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ wpart, int n_slabs, int with_rgb,
                                                           float* __restrict__ dWd, float* __restrict__ dWc) {
    __shared__ float part[8][32];
    wgrad_reduce_block(blockIdx.x, wpart, n_slabs, with_rgb, dWd, dWc, part);
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
    ARN_REQUIRE(n >= 0 && loss_scale > 0, "bad size / loss_scale");
    if (n == 0) return ARN_OK;
    ARN_REQUIRE(xyzs && params_xyz_f16 && ws.feat && ws.hid && ws.wimg && sigmas && dfeat_scratch && grad_params_xyz, "null pointer");
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
    int n_sm = 0;
    if (int e = arn::mlp_device_setup(&n_sm)) return e;
    // Pipelined like the forward: the hash-grid backward (L2 reductions) of range p runs on a second stream under the MLP
    // backward (tensor core + HBM loads) of range p+1.  Every MLP launch writes its own block of weight-gradient slabs;
    // the slab sum rides in the first hash-grid launch, otherwise it is launched on its own.
    const int64_t n_tiles = (n + 127) / 128;
    const bool runs = tunable(kTunHashBwMode) != 0;
    const int parts = (dL_dxyzs || !runs) ? 1 : pipeline_parts(n);
    // three warpgroups per CTA, one CTA per SM, when the launch has the tiles to fill that; else one warpgroup per CTA
    const bool wide = (n_tiles + parts - 1) / parts >= (int64_t)n_sm * 3 && (tunable(kTunMlpWide) & 2) != 0;
    const int wgs = wide ? 3 : 1;
    int grid = (int)max((int64_t)1, min((int64_t)n_sm * (wide ? 1 : 2), ((n_tiles + parts - 1) / parts + wgs - 1) / wgs));
    if (grid * parts > arn::kMaxBwCtas) grid = arn::kMaxBwCtas / parts;
    float* wpart = reinterpret_cast<float*>((uint8_t*)ws.wimg + arn::kWimgBytes);  // slabs follow the weight image in the scratch
    PipeStreams* ps = nullptr;
    if (parts > 1) { if (int e = pipe_streams(&ps)) return e; }
    for (int p = 0; p < parts; p++) {
#define ARN_BW_ARGS n, n_dev, dL_dsigmas, with_rgb ? dL_drgbs : nullptr, rgbs, sigmas, (const __half*)ws.feat, (const __half*)ws.hid, (const __half*)ws.in32, \
            (const __half*)ws.hid1, (const __half*)ws.hid2, (const uint8_t*)ws.wimg, rgb_act, with_rgb ? 1 : 0, loss_scale, expf(15.0f), dfeat_scratch, wpart, p, parts, p * grid, (tunable(kTunMlpWide) & 4) ? 1 : 0
        if (wide) ARN_LAUNCH_PDL("field_mlp_bw_tc_kernel", st, (arn::field_mlp_bw_tc_kernel<3>), grid, 384, arn::bw_smem_bytes<3>(), ARN_BW_ARGS);
        else ARN_LAUNCH_PDL("field_mlp_bw_tc_kernel", st, (arn::field_mlp_bw_tc_kernel<1>), grid, 128, arn::bw_smem_bytes<1>(), ARN_BW_ARGS);
#undef ARN_BW_ARGS
        if (int e = check_launch("field_mlp_bw_tc")) return e;
        if (parts > 1) ARN_CUDA(cudaEventRecord(ps->ev[p], st));
    }
    if (int e = train_fork(3, st)) return e;  // arn_train_set_fork: in front of the hash-grid backward
    WgradReduce red{wpart, grid * parts, with_rgb ? 1 : 0, grad_params_xyz, grad_params_rgb};
    if (!runs) {
        ARN_LAUNCH("wgrad_reduce_kernel", st, arn::wgrad_reduce_kernel<<<arn::kWgradFloats / 32, 256, 0, st>>>(red.wpart, red.n_slabs, red.with_rgb, red.dWd, red.dWc));
        if (int e = check_launch("wgrad_reduce")) return e;
        red.wpart = nullptr;
    }
    const WgradReduce none{nullptr, 0, 0, nullptr, nullptr};
    for (int p = 0; p < parts; p++) {
        cudaStream_t hs = parts > 1 ? ps->side : st;
        if (parts > 1) ARN_CUDA(cudaStreamWaitEvent(hs, ps->ev[p], 0));
        if (int e = hash_encode_bw_impl(xyzs, n, n_dev, xyz_min_host, xyz_max_host, levels, pxyz + ARN_DENSITY_MLP_PARAMS, dfeat_scratch,
                                        grad_params_xyz + ARN_DENSITY_MLP_PARAMS, p == parts - 1 ? dL_dxyzs : nullptr, /*tile_image=*/1, hs,
                                        p == parts - 1 ? red : none, p, parts)) return e;
    }
    if (parts > 1) {
        ARN_CUDA(cudaEventRecord(ps->join, ps->side));
        ARN_CUDA(cudaStreamWaitEvent(st, ps->join, 0));
    }
    return ARN_OK;
}
