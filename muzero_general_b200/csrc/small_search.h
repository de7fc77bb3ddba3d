// Fused search kernel for small residual networks (TicTacToe 3x3x16, the 6x6x16 hidden board of the Breakout
// configuration): ALL simulations of a search in ONE launch (small_search.cu).
#pragma once
#include <cuda_runtime.h>

#include "heads.cuh"
#include "pipeline.h"
#include "small_tower.h"

namespace mz {

struct SmallSearchArgs {
    SmallTowerArgs dyn, pred;        // dynamics tower (stem + blocks, gathered parent state + action plane) and prediction tower;
                                     // boards_per_cta = cap_channels-compatible tile of `tile` boards, n / tile loops unused
    HeadsArgs heads_dyn;             // reward head on the raw state + rescale -> hidden pool slot and dense scratch
    HeadsArgs heads_pred;            // value + policy heads
    TreeStepArgs tree;               // expand + backup + next selection (+ read-out) of a game; sim / phase flags set per simulation
    int n, g0;                       // games [g0, g0 + n) (global indices into every array)
    int n_sims;                      // simulations: tree steps sim = 1 .. n_sims (the root step ran before the launch)
    int tile;                        // games per CTA
    int first_slot;                  // pool slot receiving the state of the first simulation
    int heads_lo, heads_floats;      // slice of the head blob staged in shared memory (covers the three heads)
    int scratch_floats;              // per-warp scratch of the heads
    int off_wd, off_wp, off_wh, off_scratch, off_map, off_act;   // shared-memory layout (float offsets)
};

// P x CO: thread mapping of the towers (small_tower.cuh), G: lanes per game of the tree step (>= |A|)
bool small_search_shape(int H, int W, int C, int A, int n, int sm_count, int tower_floats, int heads_floats, int scratch_floats,
                        int cap_channels, int* P, int* CO, int* G, int* tile, int* threads, size_t* smem, int* row_stride, int* board_stride);
cudaError_t launch_small_search(SmallSearchArgs a, int P, int CO, int G, int threads, size_t smem, cudaStream_t stream);

}  // namespace mz
