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

// ---- per-kernel event timers ---------------------------------------------------------------------------
#include <map>
#include <mutex>
#include <string>
#include <vector>
namespace {
struct ProfRec {
    std::string label;
    cudaEvent_t a, b;
};
bool g_prof_on = false;
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
thread_local cudaEvent_t g_open_a = nullptr;
thread_local std::string g_open_label;
}  // namespace

void fsw_prof_begin(const char* label, cudaStream_t st) {
    if (!g_prof_on) return;
    cudaEvent_t a;
    if (cudaEventCreate(&a) != cudaSuccess) return;
    cudaEventRecord(a, st);
    g_open_a = a;
    g_open_label = label;
}

void fsw_prof_end(cudaStream_t st) {
    if (!g_prof_on || g_open_a == nullptr) return;
    cudaEvent_t b;
    if (cudaEventCreate(&b) != cudaSuccess) return;
    cudaEventRecord(b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(ProfRec{g_open_label, g_open_a, b});
    g_open_a = nullptr;
}

extern "C" int fsw_profile_enable(int on) {
    g_prof_on = on != 0;
    return FSW_OK;
}

extern "C" int64_t fsw_profile_read(char* buf, int64_t buf_bytes) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : g_prof) {
        float ms = 0.f;
        cudaEventSynchronize(r.b);
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto& e = agg[r.label];
        e.first += 1;
        e.second += ms;
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    g_prof.clear();
    std::string out;
    char line[256];
    for (auto& kv : agg) {
        snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && buf_bytes > 0) {
        size_t n = out.size() < (size_t)(buf_bytes - 1) ? out.size() : (size_t)(buf_bytes - 1);
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int64_t)out.size() + 1;
}

extern "C" int fsw_version(void) { return 100; }
extern "C" const char* fsw_last_error(void) { return g_err; }
extern "C" int fsw_built_for_sm(void) { return 100; }
extern "C" int64_t fsw_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
