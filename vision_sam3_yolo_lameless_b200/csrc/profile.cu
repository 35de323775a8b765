// Launch accounting for the C-ABI: a process-wide launch counter (bench.py's gpu_launches) and an optional
// per-launch CUDA-event profiler (cre_profile_start / cre_profile_stop) that brackets every kernel launch of
// the library with an event pair on the launching stream and reports (kernel id, milliseconds, algorithmic work).
// Off by default: the hot calls then pay one relaxed atomic increment per launch and nothing else.
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "internal.h"

namespace cre {

namespace {
std::atomic<int64_t> g_launches{0};
struct Record {
    int id;
    double work;
    cudaEvent_t e0, e1;
};
struct Profiler {
    std::mutex mu;
    bool on = false;
    size_t cap = 0;
    std::vector<Record> recs;
    std::vector<cudaEvent_t> pool;
} g_prof;

cudaEvent_t take_event() {
    if (!g_prof.pool.empty()) {
        cudaEvent_t e = g_prof.pool.back();
        g_prof.pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

LaunchScope::LaunchScope(int id, double work, cudaStream_t stream) : stream_(stream), slot_(-1) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_prof.on) return;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (!g_prof.on || g_prof.recs.size() >= g_prof.cap) return;
    Record r{id, work, take_event(), take_event()};
    if (r.e0 == nullptr || r.e1 == nullptr) return;
    cudaEventRecord(r.e0, stream);
    g_prof.recs.push_back(r);
    slot_ = static_cast<int>(g_prof.recs.size()) - 1;
}
LaunchScope::~LaunchScope() {
    if (slot_ < 0) return;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (slot_ < static_cast<int>(g_prof.recs.size())) cudaEventRecord(g_prof.recs[slot_].e1, stream_);
}

int64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

int profile_start(int max_launches) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (auto& r : g_prof.recs) {
        g_prof.pool.push_back(r.e0);
        g_prof.pool.push_back(r.e1);
    }
    g_prof.recs.clear();
    g_prof.cap = max_launches > 0 ? static_cast<size_t>(max_launches) : 0;
    g_prof.recs.reserve(g_prof.cap);
    g_prof.on = g_prof.cap > 0;
    return 0;
}

int profile_stop(int32_t* ids, float* ms, double* work, int cap) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.on = false;
    int n = 0;
    for (auto& r : g_prof.recs) {
        if (n < cap) {
            float t = 0.0f;
            if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) {
                set_error("profile_stop: event read failed: %s", cudaGetErrorString(cudaGetLastError()));
                return -2;
            }
            ids[n] = r.id;
            ms[n] = t;
            work[n] = r.work;
            ++n;
        }
        g_prof.pool.push_back(r.e0);
        g_prof.pool.push_back(r.e1);
    }
    g_prof.recs.clear();
    return n;
}

}  // namespace cre
