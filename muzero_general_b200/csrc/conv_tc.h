// tcgen05 implicit-GEMM conv3x3 on the P64C4 board layout (conv_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mz {

struct ConvTcArgs {
    const float* in;              // [n][16][64][4]  (or gathered from the hidden pool, same per-state layout)
    float* out;                   // [n][16][64][4]
    const float* residual;        // like out, or nullptr
    const float* w;               // [9][16][64][4] tf32-rounded, BN scale folded
    const float* bias;            // [64] folded BN shift, or nullptr
    const int32_t* gather_parent; // pool mode: state of game g = in + (g*pool_stride + gather_parent[g]) * 4096
    const int32_t* action;        // dynamics: add (action/A) * action_table[p][cout]
    const float* action_table;    // [64][64]
    int pool_stride;
    int n, H, W, A, relu;
    int debug_skip;               // profiling only: 1 = no MMA, 2 = no A-tile loads, 4 = no global stores, 8 = no residual loads
};

cudaError_t launch_conv3x3_tc(const ConvTcArgs& a, int sm_count, cudaStream_t stream);
bool conv_tc_supported(int C, int H, int W);
int conv_tc_board_elems();

}  // namespace mz
