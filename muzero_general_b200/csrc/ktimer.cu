#include "ktimer.h"

#include <vector>

namespace mz {
namespace {
struct Rec { int cls; cudaEvent_t a, b; };
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_free;
int g_open = -1;

cudaEvent_t get_event() {
    if (!g_free.empty()) { cudaEvent_t e = g_free.back(); g_free.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void kt_enable(bool on) { g_on = on; }
bool kt_enabled() { return g_on; }

void kt_begin(int cls, cudaStream_t stream) {
    if (!g_on) return;
    Rec r{cls, get_event(), get_event()};
    cudaEventRecord(r.a, stream);
    g_recs.push_back(r);
    g_open = (int)g_recs.size() - 1;
}

void kt_end(cudaStream_t stream) {
    if (!g_on || g_open < 0) return;
    cudaEventRecord(g_recs[g_open].b, stream);
    g_open = -1;
}

cudaError_t kt_collect(double ms[KT_CLASSES], int64_t count[KT_CLASSES]) {
    cudaError_t rc = cudaSuccess;
    for (const Rec& r : g_recs) {
        cudaError_t e = cudaEventSynchronize(r.b);
        float t = 0.0f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, r.a, r.b);
        if (e == cudaSuccess && r.cls >= 0 && r.cls < KT_CLASSES) { ms[r.cls] += t; count[r.cls] += 1; }
        else if (e != cudaSuccess) rc = e;
        g_free.push_back(r.a); g_free.push_back(r.b);
    }
    g_recs.clear();
    g_open = -1;
    return rc;
}

}  // namespace mz
