// Fully-connected MuZero networks evaluated by a group of G lanes for ONE sample.
//
// Device restatement of models.py:80-195 (MuZeroFullyConnectedNetwork) and models.py:630-642
// (mlp: Linear+ELU ... Linear+Identity) for inference:
//   representation  models.py:133-145   Linear stack, then per-sample min-max rescale
//   dynamics        models.py:147-170   [state | one_hot(action)] -> state', reward head on the
//                                       UN-normalised state', then rescale
//   prediction      models.py:128-131   policy head, value head
//   support_to_scalar models.py:645-666 fused behind the value / reward heads
// Weights live in shared memory as one blob: per Linear the transposed matrix [in][out]
// (lane o reads W[i*out+o]: consecutive lanes, consecutive banks) followed by the bias.
#pragma once
#include "common.cuh"
#include "../../include/mzb200.h"

namespace mz {

struct MlpDesc {
    int n;                               // number of Linear layers
    int in[MZ_MAX_LAYERS + 1];
    int out[MZ_MAX_LAYERS + 1];
    int w_off[MZ_MAX_LAYERS + 1];        // float offset of W^T [in][out] in the blob
    int b_off[MZ_MAX_LAYERS + 1];        // float offset of bias [out]
};

struct FcNet {
    MlpDesc rep, dyn, rew, val, pol;
    int blob_floats;
    int obs_elems, E, A, S, F;           // F = 2S+1
    int maxw;                            // widest activation vector
};

// y[o] = act(b[o] + sum_i x[i] W[i][o] (+ W[extra_row][o]))   for o striding over the lanes
template <int G>
MZ_DEVINL void linear_layer(const float* __restrict__ W, const float* __restrict__ b, int in, int out,
                            const float* x, float* y, bool elu, int extra_row) {
    const int lane = LaneGroup<G>::lane();
    for (int o = lane; o < out; o += G) {
        float acc = b[o];
        const float* w = W + o;
#pragma unroll 4
        for (int i = 0; i < in; ++i) acc = fmaf(x[i], w[i * out], acc);
        if (extra_row >= 0) acc += w[extra_row * out];
        y[o] = elu ? elu1(acc) : acc;
    }
    LaneGroup<G>::sync();
}

// Runs a whole MLP. x may be global or shared; s0/s1 are per-game ping-pong scratch (shared).
// For the dynamics net the first layer's input is [x | one_hot(action)]: `dense_in0` is the
// length of x and `extra_row` = dense_in0 + action selects the one-hot column (models.py:149-155).
// Returns the pointer holding the output (s0, s1 or `final_out` if given).
template <int G>
MZ_DEVINL float* mlp_forward(const MlpDesc& d, const float* blob, const float* x, float* s0, float* s1,
                             float* final_out, int dense_in0 = -1, int extra_row = -1) {
    const float* cur = x;
    float* dst = s0;
    for (int l = 0; l < d.n; ++l) {
        const bool last = (l == d.n - 1);
        float* y = (last && final_out) ? final_out : dst;
        const int in = (l == 0 && dense_in0 >= 0) ? dense_in0 : d.in[l];
        linear_layer<G>(blob + d.w_off[l], blob + d.b_off[l], in, d.out[l], cur, y, !last,
                        (l == 0) ? extra_row : -1);
        cur = y;
        dst = (y == s0) ? s1 : s0;
    }
    return const_cast<float*>(cur);
}

// Per-sample min-max rescale over n values held in shared memory (models.py:138-145,161-168):
// scale = max - min; if (scale < 1e-5) scale += 1e-5; out = (x - min) / scale
template <int G>
MZ_DEVINL void rescale_unit_range(const float* x, float* out, int n) {
    const int lane = LaneGroup<G>::lane();
    float lo = INFINITY, hi = -INFINITY;
    for (int i = lane; i < n; i += G) { lo = fminf(lo, x[i]); hi = fmaxf(hi, x[i]); }
    lo = -group_max_f32<G>(-lo);
    hi = group_max_f32<G>(hi);
    float sc = __fsub_rn(hi, lo);
    if (sc < 1e-5f) sc = __fadd_rn(sc, 1e-5f);
    for (int i = lane; i < n; i += G) out[i] = __fdiv_rn(__fsub_rn(x[i], lo), sc);
    LaneGroup<G>::sync();
}

// support_to_scalar (models.py:645-666) of F = 2S+1 logits held in shared memory.
template <int G>
MZ_DEVINL float support_to_scalar_group(const float* logits, int S) {
    const int lane = LaneGroup<G>::lane();
    const int F = 2 * S + 1;
    float m = -INFINITY;
    for (int i = lane; i < F; i += G) m = fmaxf(m, logits[i]);
    m = group_max_f32<G>(m);
    float den = 0.0f, num = 0.0f;
    for (int i = lane; i < F; i += G) {
        const float e = expf(logits[i] - m);
        den += e;
        num = fmaf((float)(i - S), e, num);
    }
    den = group_sum_f32<G>(den);
    num = group_sum_f32<G>(num);
    return inverse_value_transform(__fdiv_rn(num, den));
}

}  // namespace mz
