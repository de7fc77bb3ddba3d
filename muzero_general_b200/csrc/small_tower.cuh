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
    const int bufsz = a.boards_per_cta * a.board_stride;
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
//
// UW = false: item = tid, channel group fastest (the four lanes of a row share their input loads).
// UW = true ("uniform weights", P == W): the rows (board, y) of the tile are padded to a multiple of 32 and the channel group
// is the SLOW index, so every warp reads ONE weight address per tap (a broadcast: 1 shared-memory wavefront instead of 4)
// and its 32 lanes read 32 different rows; with an odd row stride and a board stride = H rows (mod 32) - the launcher's
// choice - those 32 rows fall into 32 different banks.  8 wavefronts per 36 FMAs instead of 17: the tower becomes
// FMA-bound instead of shared-memory-bound.  The arithmetic per output element is the same chain either way.
//
// Resident mode (the fused search kernel, small_search.cu): `map` != nullptr is a shared-memory table, dense element
// c * H*W + pos -> its offset inside a board's padded buffer; the input is then fetched 16 bytes at a time and placed through
// the table (no index arithmetic per element).  kTileInputStaged: buffer 0 already holds the input (the heads wrote the
// rescaled state there); kTileKeepOutput: the last layer's output stays in its shared-memory buffer (index returned) instead
// of going to a.out.  The stand-alone kernel passes map = nullptr, flags = 0.
constexpr int kTileInputStaged = 1, kTileKeepOutput = 2;

template <int P, int CO, bool UW>
__device__ __forceinline__ int small_tower_tile(const SmallTowerArgs& a, const float* s_w, float* s_act, int b0, int nbt, int tid, int nthreads,
                                                const int* map = nullptr, int flags = 0) {
    const int H = a.H, W = a.W, C = a.C;
    const int Wp = a.row_stride;
    const int plane = (H + 2) * Wp;
    const int segs = W / P;                                 // row segments: a thread owns P consecutive pixels of a row
    const int nb = a.boards_per_cta, bstride = a.board_stride;
    const int bufsz = nb * bstride;
    const int cgs = C / CO;
    const int items_per_board = cgs * H * segs;
    const int item = tid;
    int cgi, seg, y, b;
    if constexpr (UW) {
        const int rows_pad = (nb * H + 31) & ~31;
        const int r = item % rows_pad;
        cgi = item / rows_pad; seg = 0; y = r % H;
        b = (r < nb * H && cgi < cgs) ? r / H : nb;         // (nb: beyond every tile -> inactive)
    } else {
        cgi = item % cgs;
        seg = (item / cgs) % segs;
        y = (item / (cgs * segs)) % H;
        b = item / items_per_board;
    }
    const int HW = H * W;
    const int cin0 = a.layer[0].cin;
    const size_t sample_elems = (size_t)a.in_channels * HW;

    __syncthreads();                                   // previous tile fully consumed / initial fill visible
    // ---- stage the tower input (interior only) into buffer 0
    if (flags & kTileInputStaged) {
        // nothing to do
    } else if (map != nullptr) {
        const int se4 = (int)(sample_elems >> 2);                  // the launcher guarantees sample_elems % 4 == 0
        for (int i = tid; i < nbt * se4; i += nthreads) {
            const int bb = i / se4, q = i - bb * se4;
            const int g = a.g0 + b0 + bb;
            const float* src = a.gather_parent ? a.in + ((size_t)g * a.pool_stride + a.gather_parent[g]) * sample_elems
                                               : a.in + (size_t)g * sample_elems;
            const float4 v = reinterpret_cast<const float4*>(src)[q];
            float* dst = s_act + bb * bstride;
            dst[map[4 * q]] = v.x; dst[map[4 * q + 1]] = v.y; dst[map[4 * q + 2]] = v.z; dst[map[4 * q + 3]] = v.w;
        }
        if (cin0 > a.in_channels) {
            for (int i = tid; i < nbt * HW; i += nthreads) {
                const int bb = i / HW, pos = i - bb * HW;
                s_act[bb * bstride + a.in_channels * plane + map[pos]] =
                    __fdiv_rn((float)a.action[a.g0 + b0 + bb], (float)a.A);        // action / |A| plane (models.py:586-600)
            }
        }
    } else {
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
            s_act[bb * bstride + ci * plane + (yy + 1) * Wp + x + 1] = v;
        }
    }
    if (!(flags & kTileInputStaged)) __syncthreads();      // (a staged input became visible with the barrier above)

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
            const float* ib = sin + b * bstride + y * Wp + seg * P;
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
                float* so = sout + b * bstride + co * plane + (y + 1) * Wp + 1 + seg * P;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    float r = acc[c][p] + bias;
                    if (a.layer[l].residual) r += so[p];
                    if (a.layer[l].relu) r = fmaxf(r, 0.0f);
                    if (last && !(flags & kTileKeepOutput)) a.out[(((size_t)g * C + co) * H + y) * W + seg * P + p] = r;
                    else so[p] = r;
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    return cur;                                        // buffer holding the last layer's output (kTileKeepOutput)
}

}  // namespace mz
