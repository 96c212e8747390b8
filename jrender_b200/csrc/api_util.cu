// api_util.cu -- thread-local error message, process-wide launch counter, event profiler.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "../../include/b200raster.h"
#include "api_util.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
struct Span { int k; cudaEvent_t a, b; };
std::vector<Span> g_spans;          // recorded, not yet harvested
std::vector<cudaEvent_t> g_pool;    // reusable events
thread_local cudaEvent_t g_open = nullptr;
double g_total_ms[B200R_K_COUNT] = {0};
long long g_count[B200R_K_COUNT] = {0};

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void harvest_locked() {
    for (Span& s : g_spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            g_total_ms[s.k] += ms;
            g_count[s.k] += 1;
        }
        g_pool.push_back(s.a);
        g_pool.push_back(s.b);
    }
    g_spans.clear();
}
}  // namespace

int b200r_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int b200r_cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return (int)e;
}

void b200r_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }

void b200r_prof_begin(int, cudaStream_t st) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_open = get_event();
    cudaEventRecord(g_open, st);
}

void b200r_prof_end(int kernel, cudaStream_t st) {
    if (!g_prof_on.load(std::memory_order_relaxed) || g_open == nullptr) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t b = get_event();
    cudaEventRecord(b, st);
    g_spans.push_back(Span{kernel, g_open, b});
    g_open = nullptr;
}

extern "C" {
const char* b200r_last_error(void) { return g_err; }
unsigned long long b200r_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void b200r_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); }

void b200r_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    harvest_locked();
    for (int k = 0; k < B200R_K_COUNT; k++) { g_total_ms[k] = 0; g_count[k] = 0; }
}

int b200r_profile_read(int kernel, double* total_ms, long long* launches) {
    if (kernel < 0 || kernel >= B200R_K_COUNT) return b200r_fail(B200R_EINVAL, "b200r_profile_read: kernel id %d", kernel);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    harvest_locked();
    if (total_ms) *total_ms = g_total_ms[kernel];
    if (launches) *launches = g_count[kernel];
    return 0;
}
}
