// Shared host-side plumbing of libarnerf.so: error reporting, launch checks, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/arnerf.h"

#define ARN_API __attribute__((visibility("default")))

namespace arn {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Optional per-kernel device timing (bench.py): when enabled, every launch is bracketed by CUDA events on its own
// stream.  Not capturable into a CUDA graph (the caller keeps it off during capture).
struct LaunchTimer {
    LaunchTimer(const char* name, cudaStream_t st);
    ~LaunchTimer();
    const char* name_; cudaStream_t st_; void* slot_;
};

// Kernel-variant switches for A/B measurements (arn_set_tunable): every variant computes the same results.
enum Tunable { kTunMarchWarp = 0, kTunHashBwMode, kTunAdamVec, kTunPipelineParts, kTunHashBwBlocks, kTunMlpWide, kTunPdl, kTunP2pMc, kTunCount };
int tunable(Tunable t);

inline int check_launch(const char* what) {
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return ARN_E_CUDA; }
    return ARN_OK;
}

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

}  // namespace arn

#define ARN_LAUNCH(name, st, ...)                 \
    do {                                          \
        arn::LaunchTimer lt__((name), (st));      \
        __VA_ARGS__;                              \
    } while (0)

// Programmatic dependent launch for the kernels of the serial chains (training step, test-loop iteration): the launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization, every block of such a kernel starts with pdl_enter() -- wait until the
// preceding kernel of the stream has completed and its memory is visible, then allow the NEXT kernel of the stream to be
// launched -- so that the next kernel's launch latency and block scheduling overlap this kernel's execution instead of
// following it.  A kernel is launched early only once ALL blocks of its predecessor have started (each has passed its own
// wait), so nothing it holds can starve the predecessor; results are those of plain stream order.  (Captured into a CUDA
// graph the attribute becomes a programmatic dependency edge.)  "pdl" tunable, default 0 = plain launches: measured on B200 the
// early-launched blocks cost the big kernels more than the hidden launch latency returns -- training step 0.333 -> 0.365 ms,
// 800x800 frame 5.77 -> 6.42 ms; only a rank's 1/8 share of a frame gains (2.00 -> 1.93 ms).
namespace arn {
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = tunable(kTunPdl) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
}  // namespace arn
#define ARN_LAUNCH_PDL(name, st, kernel, grid, block, smem, ...)                       \
    do {                                                                               \
        arn::LaunchTimer lt__((name), (st));                                           \
        arn::launch_pdl(kernel, (grid), (block), (smem), (st), __VA_ARGS__);           \
    } while (0)

#define ARN_REQUIRE(cond, msg)                                            \
    do {                                                                  \
        if (!(cond)) { arn::set_error("%s: %s", __func__, msg); return ARN_E_INVALID; } \
    } while (0)

#define ARN_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) { arn::set_error("%s: %s", #call, cudaGetErrorString(e__)); return ARN_E_CUDA; } \
    } while (0)
