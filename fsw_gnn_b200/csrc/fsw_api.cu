// Library-level plumbing: error string, launch counter, version.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "fsw_common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

int fsw_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void fsw_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" int fsw_version(void) { return 100; }
extern "C" const char* fsw_last_error(void) { return g_err; }
extern "C" int fsw_built_for_sm(void) { return 100; }
extern "C" int64_t fsw_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
