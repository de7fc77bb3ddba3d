// Per-kernel-class device timing (CUDA events on the launching stream) for bench.py's roofline line.
// Disabled by default: kt_begin / kt_end are a single branch.  While enabled the step-wise pipeline is
// launched kernel by kernel (no CUDA-graph replay) so that every launch can be bracketed by events.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace mz {

enum KernelClass { KT_TREE = 0, KT_TOWER = 1, KT_HEADS = 2, KT_CONV = 3, KT_OTHER = 4, KT_SMALL = 5, KT_SEARCH = 6, KT_CLASSES = 7 };

void kt_enable(bool on);
bool kt_enabled();
void kt_begin(int cls, cudaStream_t stream);
void kt_end(cudaStream_t stream);
// synchronises the recorded events, adds their durations to ms[cls] / count[cls] and forgets them
cudaError_t kt_collect(double ms[KT_CLASSES], int64_t count[KT_CLASSES]);

struct KtScope {
    cudaStream_t s;
    KtScope(int cls, cudaStream_t stream) : s(stream) { kt_begin(cls, stream); }
    ~KtScope() { kt_end(s); }
};

}  // namespace mz
