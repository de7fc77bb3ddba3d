// Device code of the fused CUDA-core residual tower (small_tower.h), shared by small_tower_kernel (small_tower.cu) and the
// fused small-network search kernel (small_search.cu).  Arithmetic: fp32 FMA chain over (cin, dy, dx), + bias, + residual,
// ReLU - the order of conv3x3_kernel (resnet.cu), so every path gives bit-identical results.
//
// Thread mapping: one item = (board, row y, row segment, group of CO output channels) computes P consecutive pixels of that
// row for CO channels (P x CO accumulators).
#pragma once
#include "small_tower.h"

namespace mz {

// once per CTA: weights + biases of every layer into s_w, zeroed activation buffers (the padding stays zero afterwards)
__device__ __forceinline__ void small_tower_stage(const SmallTowerArgs& a, float* s_w, float* s_act, int tid, int nthreads) {
    const int C = a.C;
    const int plane = (a.H + 2) * (a.W + 2);
    const int bufsz = a.boards_per_cta * a.cap_channels * plane;
    for (int l = 0; l < a.n_layers; ++l) {
        const int count = a.layer[l].cin * 9 * C;
        const float4* src = reinterpret_cast<const float4*>(a.blob + a.layer[l].w_off);
        float4* dst = reinterpret_cast<float4*>(s_w + a.w_smem_off[l]);
        for (int i = tid; i < count / 4; i += nthreads) dst[i] = src[i];
        for (int i = tid; i < C; i += nthreads)
            s_w[a.b_smem_off[l] + i] = a.layer[l].b_off >= 0 ? a.blob[a.layer[l].b_off + i] : 0.0f;
    }
    for (int i = tid; i < 2 * bufsz; i += nthreads) s_act[i] = 0.0f;
}

// One tile = boards [b0, b0 + nbt) of the launch (global board index a.g0 + b0 + b) through all the layers.  Every thread of
// the CTA must call it (it synchronises the CTA); threads whose item lies beyond the tile only take part in the barriers.
template <int P, int CO>
__device__ __forceinline__ void small_tower_tile(const SmallTowerArgs& a, const float* s_w, float* s_act, int b0, int nbt, int tid, int nthreads) {
    const int H = a.H, W = a.W, C = a.C;
    const int Wp = W + 2;
    const int plane = (H + 2) * Wp;
    const int segs = W / P;                                 // row segments: a thread owns P consecutive pixels of a row
    const int nb = a.boards_per_cta, cap = a.cap_channels;
    const int bufsz = nb * cap * plane;
    const int cgs = C / CO;
    const int items_per_board = cgs * H * segs;
    const int item = tid;
    const int cgi = item % cgs;
    const int seg = (item / cgs) % segs;
    const int y = (item / (cgs * segs)) % H;
    const int b = item / items_per_board;
    const int HW = H * W;
    const int cin0 = a.layer[0].cin;
    const size_t sample_elems = (size_t)a.in_channels * HW;

    __syncthreads();                                   // previous tile fully consumed / initial fill visible
    // ---- stage the tower input (interior only) into buffer 0
    for (int i = tid; i < nbt * cin0 * HW; i += nthreads) {
        const int x = i % W, yy = (i / W) % H, ci = (i / HW) % cin0, bb = i / (HW * cin0);
        const int g = a.g0 + b0 + bb;
        float v;
        if (ci < a.in_channels) {
            const float* src = a.gather_parent ? a.in + ((size_t)g * a.pool_stride + a.gather_parent[g]) * sample_elems
                                               : a.in + (size_t)g * sample_elems;
            v = src[ci * HW + yy * W + x];
        } else {
            v = __fdiv_rn((float)a.action[g], (float)a.A);        // action / |A| plane (models.py:586-600)
        }
        s_act[(bb * cap + ci) * plane + (yy + 1) * Wp + x + 1] = v;
    }
    __syncthreads();

    const bool active = b < nbt;
    int cur = 0;
    for (int l = 0; l < a.n_layers; ++l) {
        const float* sin = s_act + cur * bufsz;
        float* sout = s_act + (cur ^ 1) * bufsz;
        if (active) {
            float acc[CO][P];
#pragma unroll
            for (int c = 0; c < CO; ++c)
#pragma unroll
                for (int p = 0; p < P; ++p) acc[c][p] = 0.0f;
            const float* ib = sin + b * cap * plane + y * Wp + seg * P;
            const float* wb = s_w + a.w_smem_off[l] + cgi * CO;
            const int cin = a.layer[l].cin;
            for (int ci = 0; ci < cin; ++ci) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    float v[P + 2];
#pragma unroll
                    for (int j = 0; j < P + 2; ++j) v[j] = ib[ci * plane + dy * Wp + j];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        float w[CO];
                        if constexpr (CO == 4) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wb + (ci * 9 + dy * 3 + dx) * C);
                            w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
                        } else {
#pragma unroll
                            for (int c = 0; c < CO; ++c) w[c] = wb[(ci * 9 + dy * 3 + dx) * C + c];
                        }
#pragma unroll
                        for (int p = 0; p < P; ++p)
#pragma unroll
                            for (int c = 0; c < CO; ++c) acc[c][p] = fmaf(v[p + dx], w[c], acc[c][p]);
                    }
                }
            }
            const bool last = l == a.n_layers - 1;
            const int g = a.g0 + b0 + b;
#pragma unroll
            for (int c = 0; c < CO; ++c) {
                const int co = cgi * CO + c;
                const float bias = s_w[a.b_smem_off[l] + co];
                float* so = sout + (b * cap + co) * plane + (y + 1) * Wp + 1 + seg * P;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float r = acc[c][p] + bias;
                    if (a.layer[l].residual) r += so[p];
                    if (a.layer[l].relu) r = fmaxf(r, 0.0f);
                    if (last) a.out[(((size_t)g * C + co) * H + y) * W + seg * P + p] = r;
                    else so[p] = r;
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

}  // namespace mz
