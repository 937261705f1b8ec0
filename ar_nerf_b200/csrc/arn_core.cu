// Error string, version and launch counter of libarnerf.so.
#include <atomic>
#include <map>
#include <mutex>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <vector>

#include "arn_common.cuh"
#include "arn_field.cuh"

namespace arn {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- kernel-variant switches
static std::atomic<int> g_tun[kTunCount] = {{1}, {16}, {1}, {1}, {0}, {3}, {0}, {4256}};
static const char* const g_tun_names[kTunCount] = {"march_warp", "hash_bw_mode", "adam_vec", "pipeline_parts", "hash_bw_blocks", "mlp_wide", "pdl", "p2p_mc"};
int tunable(Tunable t) { return g_tun[t].load(std::memory_order_relaxed); }

// ---- side stream + events for the pipelined field evaluation, one set per device
int pipe_streams(PipeStreams** out) {
    static PipeStreams pool[16];
    static bool made[16] = {};
    static std::mutex mu;
    int dev = 0;
    ARN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) { set_error("pipe_streams: device index out of range"); return ARN_E_INVALID; }
    std::lock_guard<std::mutex> lk(mu);
    if (!made[dev]) {
        PipeStreams& p = pool[dev];
        ARN_CUDA(cudaStreamCreateWithFlags(&p.side, cudaStreamNonBlocking));
        ARN_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
        ARN_CUDA(cudaEventCreateWithFlags(&p.join, cudaEventDisableTiming));
        for (int k = 0; k < 8; k++) ARN_CUDA(cudaEventCreateWithFlags(&p.ev[k], cudaEventDisableTiming));
        made[dev] = true;
    }
    *out = &pool[dev];
    return ARN_OK;
}

// ---- per-kernel timing
struct TimedLaunch { const char* name; cudaEvent_t e0, e1; };
static std::atomic<int> g_profile{0};
static std::mutex g_profile_mu;
static std::vector<TimedLaunch> g_timed;

LaunchTimer::LaunchTimer(const char* name, cudaStream_t st) : name_(name), st_(st), slot_(nullptr) {
    if (!g_profile.load(std::memory_order_relaxed)) return;
    TimedLaunch* t = new TimedLaunch{name, nullptr, nullptr};
    if (cudaEventCreate(&t->e0) != cudaSuccess || cudaEventCreate(&t->e1) != cudaSuccess) { delete t; return; }
    cudaEventRecord(t->e0, st);
    slot_ = t;
}
LaunchTimer::~LaunchTimer() {
    if (!slot_) return;
    TimedLaunch* t = (TimedLaunch*)slot_;
    cudaEventRecord(t->e1, st_);
    std::lock_guard<std::mutex> lk(g_profile_mu);
    g_timed.push_back(*t);
    delete t;
}
}  // namespace arn

extern "C" ARN_API int arn_profile_enable(int on) {
    arn::g_profile.store(on ? 1 : 0);
    return ARN_OK;
}
// Writes "kernel calls total_ms\n" lines (device time from CUDA events) for every launch since the last report, then
// clears the log.  Synchronises on the recorded events.
extern "C" ARN_API int arn_profile_report(char* buf, int cap) {
    if (!buf || cap <= 0) { arn::set_error("arn_profile_report: bad buffer"); return ARN_E_INVALID; }
    std::lock_guard<std::mutex> lk(arn::g_profile_mu);
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& t : arn::g_timed) {
        float ms = 0.f;
        if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess) {
            auto& a = agg[t.name]; a.first += 1; a.second += ms;
        }
        cudaEventDestroy(t.e0); cudaEventDestroy(t.e1);
    }
    arn::g_timed.clear();
    int off = 0; buf[0] = 0;
    for (auto& kv : agg) {
        int n = snprintf(buf + off, cap - off, "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        if (n < 0 || n >= cap - off) break;
        off += n;
    }
    return ARN_OK;
}

extern "C" ARN_API int arn_set_tunable(const char* name, int value) {
    for (int i = 0; name && i < arn::kTunCount; i++)
        if (!strcmp(name, arn::g_tun_names[i])) { arn::g_tun[i].store(value); return ARN_OK; }
    arn::set_error("arn_set_tunable: unknown tunable '%s'", name ? name : "(null)");
    return ARN_E_INVALID;
}

extern "C" ARN_API int arn_version(void) { return ARN_VERSION; }
extern "C" ARN_API const char* arn_last_error(void) { return arn::g_err; }
extern "C" ARN_API int64_t arn_launch_count(void) { return arn::g_launches.load(std::memory_order_relaxed); }
