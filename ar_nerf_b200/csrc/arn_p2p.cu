// libarnerf.so -- multi-GPU gradient exchange fused with the optimizer, over NVLink peer memory.
//
// Data-parallel training replicates the hash table; per step every rank must (1) sum the table gradients of all
// ranks, (2) run Adam, (3) see the updated fp16 table.  With NCCL that is reduce-scatter (fp32) -> Adam on the rank's
// slice -> all-gather (fp16): three serial phases, ~170 us at 8 GPUs for the 45.8 MB table.  Here ONE kernel does all
// three for the rank's slice: it LOADS the slice of every rank's gradient buffer straight from peer memory (fixed rank
// order: deterministic sum), applies Adam to the fp32 master / moments it owns, and STORES the fp16 result into every
// rank's working copy through the same peer mappings -- inbound gradient loads and outbound fp16 stores use the two
// directions of the links at once and overlap the arithmetic element by element.
//
// Ranks are separate processes (one per GPU): buffers come from arn_p2p_alloc (cudaMalloc, so that the CUDA IPC handle
// of the base pointer can be exported) and are opened by the peers with arn_p2p_open.  Synchronisation is a pair of
// monotonic flag slots per rank in peer-visible memory: arn_p2p_signal publishes "my step s is done" into every rank's
// flag array (after a system-scope fence), arn_p2p_wait spins until all ranks have published >= s.  Each rank always
// signals before it waits on the same stream, so no rank can wait for a kernel that is not already queued on its peer;
// the spin is bounded (arn_p2p_set_timeout, default 120 s -- ranks may legitimately drift apart by a checkpoint write or a
// validation pass) so that a dead peer cannot hang the GPU for ever: a wait that times out raises a flag in a host-visible
// word (arn_p2p_set_error_word) and RETURNS -- no trap, the context stays usable -- and the host raises at its next look.
#include "arn_common.cuh"
#include <string.h>

namespace arn {

constexpr int kMaxRanks = ARN_P2P_MAX_RANKS;
struct PeerF32 { const float* p[kMaxRanks]; };
struct PeerF16 { __half* p[kMaxRanks]; };
struct PeerFlags { unsigned long long* p[kMaxRanks]; };

// spin budget in clock64 ticks and the host-visible error word (pinned, mapped) of this process's device
static long long g_timeout_ticks = (long long)240e9;  // ~120 s at 2 GHz
static int g_blocks_per_sm = 8;
static unsigned int* g_err_word = nullptr;
// device-side copy of the error state: the exchange kernels look at this one (a read of the host word from every block of a
// 1184-block grid is a PCIe round trip each)
__device__ unsigned int g_err_dev = 0u;
__device__ __forceinline__ bool spin_until(const volatile unsigned long long* f, unsigned long long value, long long budget, unsigned int* err, unsigned int code) {
    const long long t0 = clock64();
    while (*f < value) {
        __nanosleep(100);
        if (clock64() - t0 > budget) {
            atomicExch(&g_err_dev, code);
            if (err) { atomicExch_system(err, code); __threadfence_system(); }
            return false;
        }
    }
    return true;
}

__global__ void p2p_signal_kernel(PeerFlags peers, int n_ranks, int rank, int slot, unsigned long long value) {
    __threadfence_system();  // everything this stream has written (also into peer memory) is visible before the flag
    const int r = threadIdx.x;
    if (r < n_ranks) {
        volatile unsigned long long* f = peers.p[r] + slot * kMaxRanks + rank;
        *f = value;
    }
    __threadfence_system();
}

__global__ void p2p_wait_kernel(const unsigned long long* __restrict__ flags, int n_ranks, int slot, unsigned long long value, long long budget,
                                unsigned int* err) {
    const int r = threadIdx.x;
    if (r < n_ranks) spin_until(flags + slot * kMaxRanks + r, value, budget, err, 0x100u | (unsigned)r);
    __threadfence_system();
}

// signal + wait in one launch (the pair always comes together on the training path)
__global__ void p2p_barrier_kernel(PeerFlags peers, const unsigned long long* __restrict__ flags, int n_ranks, int rank, int slot, unsigned long long value,
                                   long long budget, unsigned int* err) {
    __threadfence_system();
    const int r = threadIdx.x;
    if (r < n_ranks) {
        volatile unsigned long long* f = peers.p[r] + slot * kMaxRanks + rank;
        *f = value;
        spin_until(flags + slot * kMaxRanks + r, value, budget, err, 0x200u | (unsigned)r);
    }
    __threadfence_system();
}

__device__ __forceinline__ uint2 pack_half4_(const float4& a) {
    const __half2 lo = __floats2half2_rn(a.x, a.y), hi = __floats2half2_rn(a.z, a.w);
    uint2 o; o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
    return o;
}

// Elements [lo, lo + 4*n4) of the flat parameter: sum of the ranks' gradients (rank order), Adam (arn_adam_step's
// arithmetic with inv_gs = 1 / (grad_scale * world)), fp16 result to every rank.  p / m / v point at this rank's slice.
// U float4 groups per thread and turn, all their peer loads in flight together: with few peers one group per thread leaves
// the links latency-bound (2 ranks: 16 bytes per thread in flight, 81 us for 23 MB); U * (ranks - 1) >= 8 fills them.
template <int U, int R>  // R = compile-time bound of n_ranks (register arrays)
__global__ void __launch_bounds__(256) p2p_adam_exchange_kernel(PeerF32 grads, PeerF16 p16, int n_ranks, int64_t lo, int64_t n4,
                                                                float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                                float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float inv_gs) {
    // a wait in front of this kernel timed out (a peer is gone or far behind): its gradients are not final, update nothing
    if (*reinterpret_cast<const volatile unsigned int*>(&g_err_dev) != 0u) return;
    const float lr_bc1 = lr / bc1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
        float4 gr[U][R];
        float4 m4[U], v4[U], p4[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t i = i0 + u * stride;
            if (i < n4) {
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (r < n_ranks) gr[u][r] = __ldcs(reinterpret_cast<const float4*>(grads.p[r] + lo + 4 * i));  // all peer loads in flight
                m4[u] = __ldcs(m + i); v4[u] = __ldcs(v + i); p4[u] = __ldcs(p + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t i = i0 + u * stride;
            if (i >= n4) break;
            const int64_t e = lo + 4 * i;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < R; r++)
                if (r < n_ranks) { g.x += gr[u][r].x; g.y += gr[u][r].y; g.z += gr[u][r].z; g.w += gr[u][r].w; }
            float pn[4] = {p4[u].x, p4[u].y, p4[u].z, p4[u].w}, mn[4] = {m4[u].x, m4[u].y, m4[u].z, m4[u].w}, vn[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
            const float gs[4] = {g.x * inv_gs, g.y * inv_gs, g.z * inv_gs, g.w * inv_gs};
            bool touched = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (gs[k] == 0.0f && mn[k] == 0.0f && vn[k] == 0.0f) continue;  // untouched hash entry: the update is exactly zero
                touched = true;
                mn[k] = b1 * mn[k] + (1.0f - b1) * gs[k];
                vn[k] = b2 * vn[k] + (1.0f - b2) * gs[k] * gs[k];
                const float denom = sqrtf(vn[k]) / bc2_sqrt + eps;
                pn[k] = pn[k] - lr_bc1 * (mn[k] / denom);
            }
            if (!touched) continue;  // parameter unchanged: every rank's fp16 copy already holds it
            __stcs(m + i, make_float4(mn[0], mn[1], mn[2], mn[3])); __stcs(v + i, make_float4(vn[0], vn[1], vn[2], vn[3]));
            const float4 pnew = make_float4(pn[0], pn[1], pn[2], pn[3]);
            __stcs(p + i, pnew);
            const uint2 h = pack_half4_(pnew);
#pragma unroll
            for (int r = 0; r < R; r++)
                if (r < n_ranks) *reinterpret_cast<uint2*>(p16.p[r] + e) = h;
        }
    }
}

// The same exchange through the NVSwitch's multicast (NVLS): ONE multimem.ld_reduce returns the SUM of all ranks' gradients of
// a float4 group -- the switch reads the eight copies and adds them, only the reduced 16 bytes come down this rank's link
// (5.7 MB of inbound gradient traffic per step at 8 ranks instead of 40 MB) -- and ONE multimem.st writes the fp16 result into
// every rank's working copy (2.9 MB outbound instead of 20 MB).  What stays is the switch's reads of this rank's gradient
// for the other seven slices (40 MB outbound) and the other ranks' fp16 slices arriving (20 MB inbound): the links carry 43
// / 26 MB per direction instead of 60 / 60.  mc_grad / mc_p16: multicast addresses of the ranks' gradient / fp16 buffers
// (torch symmetric memory: sharding.PeerExchange).  The reduction order is the switch's (fixed by the fabric), not rank order.
template <int U>
__global__ void __launch_bounds__(U >= 8 ? 128 : (U >= 4 ? 256 : 512)) p2p_adam_exchange_mc_kernel(const float* __restrict__ mc_grad, __half* __restrict__ mc_p16, int64_t lo, int64_t n4,
                                                                   float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                                   float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float inv_gs) {
    if (*reinterpret_cast<const volatile unsigned int*>(&g_err_dev) != 0u) return;
    const float lr_bc1 = lr / bc1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
        float4 g4[U], m4[U], v4[U], p4[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t i = i0 + u * stride;
            if (i < n4) {
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(g4[u].x), "=f"(g4[u].y), "=f"(g4[u].z), "=f"(g4[u].w) : "l"(mc_grad + lo + 4 * i) : "memory");
                m4[u] = __ldcs(m + i); v4[u] = __ldcs(v + i); p4[u] = __ldcs(p + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t i = i0 + u * stride;
            if (i >= n4) break;
            const int64_t e = lo + 4 * i;
            float pn[4] = {p4[u].x, p4[u].y, p4[u].z, p4[u].w}, mn[4] = {m4[u].x, m4[u].y, m4[u].z, m4[u].w}, vn[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
            const float gs[4] = {g4[u].x * inv_gs, g4[u].y * inv_gs, g4[u].z * inv_gs, g4[u].w * inv_gs};
            bool touched = false;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (gs[k] == 0.0f && mn[k] == 0.0f && vn[k] == 0.0f) continue;  // untouched hash entry: the update is exactly zero
                touched = true;
                mn[k] = b1 * mn[k] + (1.0f - b1) * gs[k];
                vn[k] = b2 * vn[k] + (1.0f - b2) * gs[k] * gs[k];
                const float denom = sqrtf(vn[k]) / bc2_sqrt + eps;
                pn[k] = pn[k] - lr_bc1 * (mn[k] / denom);
            }
            if (!touched) continue;  // parameter unchanged: every rank's fp16 copy already holds it
            __stcs(m + i, make_float4(mn[0], mn[1], mn[2], mn[3])); __stcs(v + i, make_float4(vn[0], vn[1], vn[2], vn[3]));
            const float4 pnew = make_float4(pn[0], pn[1], pn[2], pn[3]);
            __stcs(p + i, pnew);
            const uint2 h = pack_half4_(pnew);
            asm volatile("multimem.st.relaxed.sys.global.v2.f16x2 [%0], {%1, %2};" :: "l"(mc_p16 + e), "r"(h.x), "r"(h.y) : "memory");
        }
    }
}

}  // namespace arn

using namespace arn;

extern "C" ARN_API int arn_p2p_alloc(void** ptr_host, int64_t bytes) {
    ARN_REQUIRE(ptr_host && bytes > 0, "bad arguments");
    ARN_CUDA(cudaMalloc(ptr_host, (size_t)bytes));
    ARN_CUDA(cudaMemset(*ptr_host, 0, (size_t)bytes));
    return ARN_OK;
}
extern "C" ARN_API int arn_p2p_free(void* ptr) {
    if (ptr) ARN_CUDA(cudaFree(ptr));
    return ARN_OK;
}
extern "C" ARN_API int arn_p2p_export(void* ptr, unsigned char* handle64_host) {
    ARN_REQUIRE(ptr && handle64_host, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    ARN_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64_host, &h, 64);
    return ARN_OK;
}
extern "C" ARN_API int arn_p2p_open(const unsigned char* handle64_host, void** ptr_host) {
    ARN_REQUIRE(handle64_host && ptr_host, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    ARN_CUDA(cudaIpcOpenMemHandle(ptr_host, h, cudaIpcMemLazyEnablePeerAccess));
    return ARN_OK;
}
extern "C" ARN_API int arn_p2p_close(void* ptr) {
    if (ptr) ARN_CUDA(cudaIpcCloseMemHandle(ptr));
    return ARN_OK;
}

extern "C" ARN_API int arn_p2p_signal(void* const* peer_flags_host, int n_ranks, int rank, int slot, uint64_t value, arn_stream_t stream) {
    ARN_REQUIRE(peer_flags_host && n_ranks >= 1 && n_ranks <= kMaxRanks && rank >= 0 && rank < n_ranks && slot >= 0 && slot < ARN_P2P_FLAG_SLOTS, "bad arguments");
    PeerFlags pf{};
    for (int r = 0; r < n_ranks; r++) { ARN_REQUIRE(peer_flags_host[r], "null peer flag array"); pf.p[r] = (unsigned long long*)peer_flags_host[r]; }
    ARN_LAUNCH("p2p_signal_kernel", (cudaStream_t)stream, p2p_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, n_ranks, rank, slot, (unsigned long long)value));
    return check_launch("p2p_signal");
}
extern "C" ARN_API int arn_p2p_wait(const void* my_flags, int n_ranks, int slot, uint64_t value, arn_stream_t stream) {
    ARN_REQUIRE(my_flags && n_ranks >= 1 && n_ranks <= kMaxRanks && slot >= 0 && slot < ARN_P2P_FLAG_SLOTS, "bad arguments");
    ARN_LAUNCH("p2p_wait_kernel", (cudaStream_t)stream, p2p_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const unsigned long long*)my_flags, n_ranks, slot, (unsigned long long)value,
                                                                                                       g_timeout_ticks, g_err_word));
    return check_launch("p2p_wait");
}

extern "C" ARN_API int arn_p2p_barrier(void* const* peer_flags_host, const void* my_flags, int n_ranks, int rank, int slot, uint64_t value, arn_stream_t stream) {
    ARN_REQUIRE(peer_flags_host && my_flags && n_ranks >= 1 && n_ranks <= kMaxRanks && rank >= 0 && rank < n_ranks && slot >= 0 && slot < ARN_P2P_FLAG_SLOTS, "bad arguments");
    PeerFlags pf{};
    for (int r = 0; r < n_ranks; r++) { ARN_REQUIRE(peer_flags_host[r], "null peer flag array"); pf.p[r] = (unsigned long long*)peer_flags_host[r]; }
    ARN_LAUNCH("p2p_barrier_kernel", (cudaStream_t)stream, p2p_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, (const unsigned long long*)my_flags, n_ranks, rank, slot, (unsigned long long)value,
                                                                                                             g_timeout_ticks, g_err_word));
    return check_launch("p2p_barrier");
}

extern "C" ARN_API int arn_p2p_adam_exchange(void* const* peer_grads_host, void* const* peer_p16_host, int n_ranks, int64_t lo, int64_t count,
                                             float* params_slice, float* exp_avg_slice, float* exp_avg_sq_slice, float lr, float beta1, float beta2,
                                             float eps, int step, float inv_grad_scale, arn_stream_t stream) {
    ARN_REQUIRE(peer_grads_host && peer_p16_host && n_ranks >= 1 && n_ranks <= kMaxRanks && step >= 1, "bad arguments");
    ARN_REQUIRE(lo >= 0 && count >= 0 && lo % 4 == 0 && count % 4 == 0, "slice must be 4-element aligned");
    if (count == 0) return ARN_OK;
    ARN_REQUIRE(params_slice && exp_avg_slice && exp_avg_sq_slice, "null pointer");
    PeerF32 g{}; PeerF16 h{};
    for (int r = 0; r < n_ranks; r++) {
        ARN_REQUIRE(peer_grads_host[r] && peer_p16_host[r], "null peer buffer");
        ARN_REQUIRE(((uintptr_t)peer_grads_host[r] & 15) == 0 && ((uintptr_t)peer_p16_host[r] & 7) == 0, "peer buffers must be 16-byte aligned");
        g.p[r] = (const float*)peer_grads_host[r]; h.p[r] = (__half*)peer_p16_host[r];
    }
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    const int64_t n4 = count / 4;
    // grid: arn_p2p_set_grid blocks per SM (default 8 = whatever fits); a caller that runs the exchange of one level group
    // beside the hash-grid backward of the next asks for fewer so that both kernels hold SM slots at the same time
    // few ranks: several float4 groups per thread (the unrolled kernels hold their loads in ~140 registers: 128-thread blocks keep
    // three of them on an SM)
    const int threads = n_ranks <= 4 ? 128 : 256;
    const int grid = (int)min((int64_t)148 * g_blocks_per_sm * (256 / threads), (n4 + threads - 1) / threads);
    cudaStream_t st = (cudaStream_t)stream;
#define ARN_P2P_EX(U, R) ARN_LAUNCH("p2p_adam_exchange_kernel", st, (p2p_adam_exchange_kernel<U, R><<<grid, threads, 0, st>>>(g, h, n_ranks, lo, n4, (float4*)params_slice, \
        (float4*)exp_avg_slice, (float4*)exp_avg_sq_slice, lr, beta1, beta2, eps, bc1, bc2_sqrt, inv_grad_scale)))
    if (n_ranks <= 2) ARN_P2P_EX(4, 2); else if (n_ranks <= 4) ARN_P2P_EX(2, 4); else ARN_P2P_EX(1, 8);
#undef ARN_P2P_EX
    return check_launch("p2p_adam_exchange");
}

extern "C" ARN_API int arn_p2p_set_grid(int blocks_per_sm) {
    ARN_REQUIRE(blocks_per_sm >= 1 && blocks_per_sm <= 16, "1..16 blocks per SM");
    g_blocks_per_sm = blocks_per_sm;
    return ARN_OK;
}
extern "C" ARN_API int arn_p2p_set_timeout(double seconds) {
    ARN_REQUIRE(seconds > 0, "timeout must be positive");
    g_timeout_ticks = (long long)(seconds * 2.0e9);
    return ARN_OK;
}
// err_word: device-accessible address of a zeroed 32-bit word in pinned, mapped host memory (or NULL: timeouts go unreported)
extern "C" ARN_API int arn_p2p_set_error_word(void* err_word) {
    g_err_word = (unsigned int*)err_word;
    return ARN_OK;
}


extern "C" ARN_API int arn_p2p_adam_exchange_mc(const void* mc_grads, void* mc_p16, int64_t lo, int64_t count, float* params_slice, float* exp_avg_slice,
                                                float* exp_avg_sq_slice, float lr, float beta1, float beta2, float eps, int step, float inv_grad_scale,
                                                arn_stream_t stream) {
    ARN_REQUIRE(mc_grads && mc_p16 && step >= 1, "bad arguments");
    ARN_REQUIRE(lo >= 0 && count >= 0 && lo % 4 == 0 && count % 4 == 0, "slice must be 4-element aligned");
    ARN_REQUIRE(((uintptr_t)mc_grads & 15) == 0 && ((uintptr_t)mc_p16 & 7) == 0, "multicast buffers must be 16-byte aligned");
    if (count == 0) return ARN_OK;
    ARN_REQUIRE(params_slice && exp_avg_slice && exp_avg_sq_slice, "null pointer");
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    const int64_t n4 = count / 4;
    // "p2p_mc" tunable = 1000 * (float4 groups per thread) + threads per block (A/B: how many multimem.ld_reduce are in flight)
    const int var = tunable(kTunP2pMc), U = var / 1000, threads = var % 1000;
    ARN_REQUIRE((U == 1 || U == 2 || U == 4 || U == 8) && (threads == 128 || threads == 256 || threads == 512), "bad p2p_mc variant");
    ARN_REQUIRE(threads <= (U >= 8 ? 128 : (U >= 4 ? 256 : 512)), "p2p_mc: too many threads for that many groups per thread");
    const int grid = (int)min((int64_t)148 * g_blocks_per_sm * (256 / min(threads, 256)), (n4 + threads - 1) / threads);
    cudaStream_t st = (cudaStream_t)stream;
#define ARN_MC(UU) ARN_LAUNCH("p2p_adam_exchange_mc_kernel", st, (p2p_adam_exchange_mc_kernel<UU><<<grid, threads, 0, st>>>((const float*)mc_grads, (__half*)mc_p16, lo, n4, \
        (float4*)params_slice, (float4*)exp_avg_slice, (float4*)exp_avg_sq_slice, lr, beta1, beta2, eps, bc1, bc2_sqrt, inv_grad_scale)))
    if (U == 1) ARN_MC(1); else if (U == 2) ARN_MC(2); else if (U == 4) ARN_MC(4); else ARN_MC(8);
#undef ARN_MC
    return check_launch("p2p_adam_exchange_mc");
}
