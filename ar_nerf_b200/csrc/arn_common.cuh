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
enum Tunable { kTunMarchWarp = 0, kTunHashBwMode, kTunAdamVec, kTunPipelineParts, kTunHashBwBlocks, kTunMlpWide, kTunCount };
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

#define ARN_REQUIRE(cond, msg)                                            \
    do {                                                                  \
        if (!(cond)) { arn::set_error("%s: %s", __func__, msg); return ARN_E_INVALID; } \
    } while (0)

#define ARN_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) { arn::set_error("%s: %s", #call, cudaGetErrorString(e__)); return ARN_E_CUDA; } \
    } while (0)
