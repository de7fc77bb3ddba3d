// The library handle and the helpers shared by the entry-point files (abi.cu, selfplay.cu).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "kernels.h"
#include "pipeline.h"
#include "ktimer.h"

struct MzSelfPlay;                     // device-resident self-play state (selfplay.cu)
using namespace mz;

struct MzHandle {
    MzNetDesc net;
    MzSearchDesc search;
    int device = 0;
    int sm_count = 0;
    size_t smem_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    int64_t launches = 0;
    double last_ms = 0.0;
    // tables
    double* d_pbc = nullptr;
    double* d_sqrt = nullptr;
    double* d_ucb = nullptr;           // optional host-evaluated exploration-factor table
    // fully-connected weights
    FcNet fc{};
    float* d_fc_blob = nullptr;
    bool weights_loaded = false;
    int fc_group = 16;
    int fc_threads = 64;
    // residual weights + workspace
    ResNetDevice* res = nullptr;
    // pool
    NodePool pool{};
    int64_t hidden_elems = 0, obs_elems = 0;
    int64_t pool_state_elems = 0;      // floats per hidden state as stored in the pool (layout dependent)
    // IO arenas
    unsigned char* d_in = nullptr;
    unsigned char* d_out = nullptr;
    unsigned char* h_in = nullptr;
    unsigned char* h_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    // CUDA graphs of the step-wise pipeline, one per argument set (callers that rotate a few input buffers - double
    // buffering, bench.py's four batches - keep replaying): a set seen twice is captured, the least recently used of
    // kMaxGraphs entries makes room
    struct SearchGraph {
        uint64_t key = 0;
        int seen = 0;
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;
        int parts = 1;
        uint64_t used = 0;             // tick of the last use
    };
    static constexpr int kMaxGraphs = 8;
    std::vector<SearchGraph> graphs;
    uint64_t graph_tick = 0;
    // partitioned replay: the simulations of disjoint game ranges run as parallel branches of the graph (abi.cu)
    static constexpr int kMaxParts = 4;
    cudaStream_t part_stream[kMaxParts] = {nullptr, nullptr, nullptr, nullptr};     // [0] unused (= stream)
    cudaEvent_t part_fork = nullptr, part_join[kMaxParts] = {nullptr, nullptr, nullptr, nullptr};
    int graph_parts = 1;               // branches of the graph replayed last (mz_graph_partitions)
    // lazily allocated debug buffers
    std::vector<void*> debug_allocs;
    std::map<std::string, std::pair<void*, size_t>> named;
    MzSelfPlay* sp = nullptr;          // mz_selfplay_begin
    int pool_n = 0;                    // layout "N" of the node pool and tables: num_simulations + extra_expansions
    int imported_expansions = 0;       // expansions of the tree mz_import_tree seeded last (MZ_FLAG_CONTINUE)
    int range_fallbacks = 0;           // times the x3 range guard switched this handle to the fp32 towers (0 or 1)
};

int mz_fail(MzHandle* h, int code, const std::string& msg);
static inline int fail(MzHandle* h, int code, const std::string& msg) { return mz_fail(h, code, msg); }

#define MZ_CUDA(h, expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(h, MZ_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));          \
    } while (0)


// One batched search on device buffers, enqueued on h->stream without synchronising: the fused FC kernel or the
// step-wise pipeline (eager for the first two calls with a given argument set, then a CUDA-graph replay).
int mz_dispatch_search(MzHandle* h, const mz::SearchCall& call, bool teacher, bool trace, int flags);
void mz_selfplay_destroy(MzHandle* h);
void mz_switch_to_strict(MzHandle* h);
void mz_drop_graphs(MzHandle* h);             // captured graphs refer to buffers or kernels that are about to change
