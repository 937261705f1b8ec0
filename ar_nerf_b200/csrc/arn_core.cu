// Error string, version and launch counter of libarnerf.so.
#include <atomic>
#include <stdarg.h>

#include "arn_common.cuh"

namespace arn {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace arn

extern "C" ARN_API int arn_version(void) { return ARN_VERSION; }
extern "C" ARN_API const char* arn_last_error(void) { return arn::g_err; }
extern "C" ARN_API int64_t arn_launch_count(void) { return arn::g_launches.load(std::memory_order_relaxed); }
