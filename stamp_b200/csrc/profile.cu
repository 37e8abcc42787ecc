// Event-pair profiler behind stamp_b200_profile_* (see profile.cuh).
#include "profile.cuh"

#include <mutex>
#include <vector>

#include "stamp_b200.h"

namespace sb {
namespace {
struct Rec { cudaEvent_t a, b; int cat; double work; };
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t g_open[PROF_NCAT];

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace

bool prof_enabled() { return g_on; }

void prof_begin(int cat, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEvent_t e = get_event();
    cudaEventRecord(e, stream);
    g_open[cat] = e;
}

void prof_end(int cat, double work, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEvent_t e = get_event();
    cudaEventRecord(e, stream);
    g_recs.push_back(Rec{g_open[cat], e, cat, work});
}
}  // namespace sb

extern "C" {

void stamp_b200_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(sb::g_mu);
    sb::g_on = on != 0;
}

// Synchronises, then fills per-category totals: ms[c], work[c], count[c] for c < ncat (<= 5):
// 0 GEMM (work = FLOPs), 1 attention (FLOPs), 2 row ops (bytes), 3 Macenko (bytes), 4 pooling (bytes).
// Clears the records. Returns the number of launches summarised.
int stamp_b200_profile_summary(double* ms, double* work, long long* count, int ncat) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(sb::g_mu);
    for (int c = 0; c < ncat; ++c) { ms[c] = 0; work[c] = 0; count[c] = 0; }
    int n = 0;
    for (auto& r : sb::g_recs) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        if (r.cat < ncat) { ms[r.cat] += t; work[r.cat] += r.work; count[r.cat] += 1; }
        sb::g_pool.push_back(r.a);
        sb::g_pool.push_back(r.b);
        ++n;
    }
    sb::g_recs.clear();
    return n;
}

}  // extern "C"
