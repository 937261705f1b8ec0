// sm_100a primitives used by the tensor-core MLP kernels: tcgen05 (UMMA) issue / commit / TMEM load, TMEM allocation,
// mbarriers, bulk (TMA) copies, and the shared-memory operand layouts.
//
// Operand layout (one rule for everything): a tile is [rows][W halves] with row bytes RB = 2W in {32, 64, 128}, stored
// densely from a 1024-byte aligned base with the hardware swizzle of the same width (Swizzle<log2(RB/16),4,3> on the
// byte offset).  The SAME bytes are
//   * a K-major UMMA operand with (M or N) = rows and K = the W columns   (SBO = 8*RB, 32 B per K=16 step), and
//   * an MN-major UMMA operand with (M or N) = the W columns and K = rows (SBO = 8*RB, 16 rows = 2*SBO per K=16 step),
// which is what lets the backward reuse one activation tile for dgrad (K-major) and wgrad (MN-major), and one weight
// image for forward (K-major) and dgrad (MN-major).  Canonical layouts: CUTLASS cute/atom/mma_traits_sm100.hpp.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace arn {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- swizzled byte offset of 16-byte chunk `c` of row `r` in a tile with RB-byte rows
template <int RB>
__device__ __forceinline__ uint32_t swz(uint32_t r, uint32_t c) {
    static_assert(RB == 32 || RB == 64 || RB == 128, "row bytes");
    const uint32_t off = r * RB + c * 16;
    constexpr uint32_t mask = RB / 16 - 1;  // 1, 3, 7
    return off ^ (((off >> 7) & mask) << 4);
}

// ---- descriptors
template <int RB>
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    constexpr uint64_t layout = RB == 128 ? 2 : (RB == 64 ? 4 : 6);  // SWIZZLE_128B / 64B / 32B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);               // start address  [0,14)
    d |= (uint64_t)1 << 16;                                // LBO (16 B units) [16,30): 1, as CUTLASS sets it for swizzled modes
    d |= (uint64_t)((8 * RB) >> 4) << 32;                  // SBO            [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= layout << 61;                                     // layout type    [61,64)
    return d;
}
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// kind::f16, fp16 x fp16 -> fp32.  a_mn / b_mn: operand is MN-major (1) or K-major (0).
__host__ __device__ constexpr uint32_t instr_desc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- tcgen05
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 columns of this warp's 32 lanes <- one value (zero-initialising accumulators several issuing threads add into)
__device__ __forceinline__ void tmem_st16_fill(uint32_t taddr, uint32_t v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n"
        :: "r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// barrier among the 128 threads of one warpgroup (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void wg_sync(int id) { asm volatile("bar.sync %0, 128;" :: "r"(id) : "memory"); }

// Whole warp.  Writes the TMEM base address to *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
// Bounded spin: a lost arrival traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();
    }
}
// Bulk global -> shared copy (TMA engine, UBLKCP in SASS); completion is signalled on mbar as a byte count.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// Bulk shared -> global copy (TMA engine); grouped, the issuing thread waits for the group's shared-memory reads.
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }

}  // namespace tc

// Weight image: the five MLP matrices as swizzled operand tiles, built by pack_mlp_weights_kernel from the fp16 params.
//   [0)      density W1 [64][32]  RB 64     4096 B        [4096)   density W2 [16][64]  RB 128   2048 B
//   [6144)   colour  W1 [64][32]  RB 64     4096 B        [10240)  colour  W2 [64][64]  RB 128   8192 B
//   [18432)  colour  W3 [16][64]  RB 128    2048 B        total 20480 B
constexpr int kWimgD1 = 0, kWimgD2 = 4096, kWimgC1 = 6144, kWimgC2 = 10240, kWimgC3 = 18432, kWimgBytes = 20480;
constexpr int kWimgChunks = 1280;  // 16-byte chunks of the image

// fp16 params (row-major [out][in] per layer, tcnn order) -> swizzled operand image: one 16-byte chunk (8 halves)
__device__ __forceinline__ void pack_weight_chunk(int chunk, const __half* __restrict__ Wd, const __half* __restrict__ Wc, uint8_t* __restrict__ img) {
    using tc::swz;
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

// Sum of the backward CTAs' weight-gradient slabs into the gradient buffers, fixed order (deterministic).  A block of 256
// threads owns 32 consecutive weights: warp w adds the slabs k = w, w+8, ... (128-byte coalesced rows), the eight partial
// sums are combined in warp order.  `part` is an 8 x 32 shared-memory scratch.
constexpr int kWgradFloats = 7168 + 3072;  // one slab: colour 7168 | density 3072
__device__ __forceinline__ void wgrad_reduce_block(int block, const float* __restrict__ wpart, int n_slabs, int with_rgb,
                                                   float* __restrict__ dWd, float* __restrict__ dWc, float (*part)[32]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int e = block * 32 + lane;
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
