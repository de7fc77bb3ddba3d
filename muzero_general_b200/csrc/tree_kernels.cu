// Step-wise tree kernels over the HBM node pool (game-major SoA, see kernels.h / tree.cuh).
//
// One launch per simulation performs, for every game of the batch,
//   [root expansion]  ->  [expand + backup of the previous simulation's leaf]  ->
//   [selection of the next leaf]  ->  [final read-out]
// (phases enabled by flags), so that a search of N simulations is N+1 tree launches with the
// batched network evaluation in between.  A group of G >= |A| lanes owns one game: the A child
// slots of a node are read with one coalesced access per array, scores are reduced with
// shuffles, and the per-node backup updates are spread over the lanes (tree.cuh).
#include "kernels.h"
#include "pipeline.h"
#include "tree.cuh"
#include "launch.h"

#include <stdlib.h>

namespace mz {

// kLatency: few games in flight (the launch is bound by the chain of dependent round trips to L2, not by
// throughput): selection uses the single-round-trip + L1-prefetch variant of tree_select (tree.cuh).
template <int G, bool kLatency>
__global__ void __launch_bounds__(128) tree_step_kernel(const __grid_constant__ TreeStepArgs a) {
    pdl_launch_dependents();
    pdl_wait();                                   // everything below reads what the network kernels just wrote
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (g >= a.n) return;
    const int lane = LaneGroup<G>::lane();
    const int N = a.N, A = a.A;
    const size_t slots = (size_t)(N + 1) * A;
    const NodePool& p = a.pool;

    TreeConst c;
    c.A = A; c.N = N; c.P = a.P; c.discount = a.discount; c.noise_frac = a.noise_frac; c.noise_alpha = a.noise_alpha; c.seed = a.seed;
    c.pbc = a.pbc; c.sqrtn = a.sqrtn; c.ucb = a.ucb;

    GameTree t;
    t.visit = p.visit + g * slots;
    t.vsum = p.vsum + g * slots;
    t.mval = p.mval + g * slots;
    t.reward = p.reward + g * slots;
    t.prior = p.prior + g * slots;
    t.expansion = p.expansion + g * slots;
    t.root_prior = p.root_prior + (size_t)g * A;
    t.path = p.path + (size_t)g * (N + 2);
    t.path_reward = p.path_reward + (size_t)g * (N + 2);
    int max_depth = 0;

    if (a.do_root == 1) {
        unsigned legal = 0;
        for (int k = 0; k < A; ++k)
            legal |= (a.legal_mask == nullptr || a.legal_mask[(size_t)g * A + k]) ? (1u << k) : 0u;
        t.legal = legal;
        const bool ok = lane < A && ((legal >> lane) & 1u);
        float prior;
        if (a.policy_is_prior) prior = (lane < A) ? a.net_policy[(size_t)g * a.policy_stride + lane] : 0.0f;
        else prior = group_softmax_masked<G>((lane < A) ? a.net_policy[(size_t)g * a.policy_stride + lane] : 0.0f, ok);
        const float root_reward = a.net_reward ? a.net_reward[(size_t)g * a.value_stride] : inverse_value_transform(0.0f);
        if (a.trace.root_priors_raw && lane < A) a.trace.root_priors_raw[(size_t)g * A + lane] = ok ? prior : 0.0f;
        if (a.trace.root_reward && lane == 0) a.trace.root_reward[g] = root_reward;
        tree_init_root<G>(c, t, prior, root_reward, (a.add_noise && a.noise) ? a.noise + (size_t)g * A : nullptr,
                          a.add_noise && !a.noise, a.game_id ? a.game_id[g] : (int64_t)g, a.move_index ? a.move_index[g] : 0,
                          a.trace.noise ? a.trace.noise + (size_t)g * A : nullptr);
        if (lane == 0 && a.root_predicted_value) a.root_predicted_value[g] = a.net_value[(size_t)g * a.value_stride];
    } else {
        t.legal = p.legal[g];
        t.root_visit = p.root_visit[g];
        t.root_vsum = p.root_vsum[g];
        t.root_reward = p.root_reward[g];
        t.lo = p.range[2 * g];
        t.hi = p.range[2 * g + 1];
        t.n_expanded = p.n_expanded[g];
        t.ties = p.ties[g];
        max_depth = p.max_depth[g];
    }

    if (a.do_update) {
        Leaf leaf;
        leaf.depth = p.leaf_depth[g];
        leaf.parent_exp = p.leaf_parent[g];
        leaf.action = p.leaf_action[g];
        leaf.slot = p.leaf_slot[g];
        const float value = a.net_value[(size_t)g * a.value_stride];
        const float reward = a.net_reward[(size_t)g * a.value_stride];
        float prior;
        if (a.policy_is_prior) prior = (lane < A) ? a.net_policy[(size_t)g * a.policy_stride + lane] : 0.0f;
        else prior = group_softmax_masked<G>((lane < A) ? a.net_policy[(size_t)g * a.policy_stride + lane] : 0.0f, lane < A);
        if (a.trace.depth) {
            const int sim = a.sim - 1;
            const size_t ti = (size_t)g * N + sim;
            if (lane == 0) { a.trace.depth[ti] = leaf.depth; a.trace.value[ti] = value; a.trace.reward[ti] = reward; }
            if (lane < A) a.trace.priors[ti * A + lane] = prior;
            for (int j = lane; j < leaf.depth && j < a.trace.max_depth; j += G)
                a.trace.actions[ti * a.trace.max_depth + j] = (uint8_t)(t.path[j + 1] % A);
        }
        tree_expand<G>(c, t, leaf, reward, prior);
        tree_backup<G>(c, t, leaf, value);
        max_depth = max(max_depth, leaf.depth);
    }

    if (a.do_select) {
        const int64_t game_id = a.game_id ? a.game_id[g] : (int64_t)g;
        const int move = a.move_index ? a.move_index[g] : 0;
        const int first_index = a.first_index ? a.first_index[g] : -1;
        const Leaf leaf = tree_select<G, kLatency>(c, t, a.sim, game_id, move, first_index);
        if (lane == 0) {
            p.leaf_depth[g] = leaf.depth;
            p.leaf_parent[g] = leaf.parent_exp;
            p.leaf_action[g] = leaf.action;
            p.leaf_slot[g] = leaf.slot;
        }
    }

    if (lane == 0) {
        p.legal[g] = t.legal;
        p.root_visit[g] = t.root_visit;
        p.root_vsum[g] = t.root_vsum;
        p.root_reward[g] = t.root_reward;
        p.range[2 * g] = t.lo;
        p.range[2 * g + 1] = t.hi;
        p.n_expanded[g] = t.n_expanded;
        p.ties[g] = t.ties;
        p.max_depth[g] = max_depth;
    }

    if (a.do_final) {
        if (lane < A) {
            const bool ok = (t.legal >> lane) & 1u;
            if (a.visit_counts) a.visit_counts[(size_t)g * A + lane] = ok ? t.visit[lane] : 0;
            if (a.root_priors) a.root_priors[(size_t)g * A + lane] = t.root_prior[lane];
        }
        if (lane == 0) {
            if (a.root_value) a.root_value[g] = (t.root_visit == 0) ? 0.0 : __ddiv_rn(t.root_vsum, (double)t.root_visit);
            if (a.max_tree_depth) a.max_tree_depth[g] = max_depth;
            if (a.tie_count) a.tie_count[g] = t.ties;
            if (a.value_range) { a.value_range[2 * g] = t.lo; a.value_range[2 * g + 1] = t.hi; }
        }
    }
}

// override_root_with (self_play.py:275-277, 310-314): the tree mz_import_tree put into the pool becomes the root of a new
// search - fresh MinMaxStats, tie / depth counters reset, and the exploration noise mixed into the priors the root's
// children already have (self_play.py:467-476).  Its own kernel (one game per lane group, launched once per continued
// search) so that the per-simulation kernel keeps its register budget.
template <int G>
__global__ void __launch_bounds__(128) tree_adopt_root_kernel(const __grid_constant__ TreeStepArgs a) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    if (g >= a.n) return;
    const int lane = LaneGroup<G>::lane();
    const int A = a.A;
    const NodePool& p = a.pool;
    if (lane == 0) {
        p.range[2 * g] = INFINITY;
        p.range[2 * g + 1] = -INFINITY;
        p.ties[g] = 0;
        p.max_depth[g] = 0;
    }
    if (a.add_noise) {
        const int64_t gid = a.game_id ? a.game_id[g] : (int64_t)g;
        const int mv = a.move_index ? a.move_index[g] : 0;
        double sum = 0.0;
        if (!a.noise) {                               // drawn here: normalised Gamma(alpha) draws over the whole action space
            for (int k = lane; k < A; k += G) sum += philox_gamma(a.seed, gid, mv, k, a.noise_alpha);
            const unsigned m = LaneGroup<G>::mask();
            for (int off = G >> 1; off > 0; off >>= 1) sum += shfl_xor_f64(m, sum, off, G);
        }
        for (int k = lane; k < A; k += G) {           // lane l owns the actions l, l + G, ...
            const double nz = a.noise ? a.noise[(size_t)g * A + k] : philox_gamma(a.seed, gid, mv, k, a.noise_alpha) / sum;
            if (a.trace.noise) a.trace.noise[(size_t)g * A + k] = nz;
            double* rp = p.root_prior + (size_t)g * A + k;
            *rp = __dadd_rn(__dmul_rn(*rp, __dsub_rn(1.0, a.noise_frac)), __dmul_rn(nz, a.noise_frac));
        }
    }
}

cudaError_t launch_tree_adopt_root(const TreeStepArgs& a, cudaStream_t stream) {
    int G = 4;
    while (G < a.A && G < 32) G <<= 1;
    const int grid = (a.n * G + 127) / 128;
    switch (G) {
        case 4: tree_adopt_root_kernel<4><<<grid, 128, 0, stream>>>(a); break;
        case 8: tree_adopt_root_kernel<8><<<grid, 128, 0, stream>>>(a); break;
        case 16: tree_adopt_root_kernel<16><<<grid, 128, 0, stream>>>(a); break;
        case 32: tree_adopt_root_kernel<32><<<grid, 128, 0, stream>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_tree_step(const TreeStepArgs& a, cudaStream_t stream) {
    if (a.A > 32) return launch_tree_step_wide(a, stream);
    int G = 4;
    while (G < a.A) G <<= 1;
    const int threads = 128;
    const int games_per_cta = threads / G;
    const int grid = (a.n + games_per_cta - 1) / games_per_cta;
    // below ~4 resident warps per scheduler the kernel is latency-bound (MZ_TREE_LATENCY = 0 / 1 forces a variant: A/B switch)
    static const int forced = getenv("MZ_TREE_LATENCY") ? atoi(getenv("MZ_TREE_LATENCY")) : -1;
    const bool latency = forced >= 0 ? forced != 0 : (long)a.n * G <= 148L * 512;
    cudaError_t e = cudaSuccess;
#define MZ_TREE(GG)                                                                                 \
    case GG:                                                                                        \
        e = latency ? launch_chained(tree_step_kernel<GG, true>, dim3(grid), dim3(threads), 0, stream, a)      \
                    : launch_chained(tree_step_kernel<GG, false>, dim3(grid), dim3(threads), 0, stream, a);    \
        break;
    switch (G) {
        MZ_TREE(4) MZ_TREE(8) MZ_TREE(16) MZ_TREE(32)
        default: return cudaErrorInvalidValue;
    }
#undef MZ_TREE
    return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace mz
