// Residual MuZero networks (models.py:206-623) on the device - fp32 CUDA-core path.
//
//   representation  models.py:339-349 (+ DownSample models.py:233-275), rescale models.py:526-553
//   dynamics        models.py:379-389, action plane + rescale models.py:555-599
//   prediction      models.py:424-433
//   support_to_scalar models.py:645-666 fused behind the value / reward heads
//
// Layout: activations NCHW fp32 in HBM workspaces; BatchNorm (eval mode, self_play.py:29) is
// folded into the conv weights/bias at load time; the dynamics input plane action/|A|
// (models.py:557-572) is synthesised while staging the input tile, never materialised.
// conv3x3 is a register-tiled direct convolution: a CTA stages the (padded) input planes of
// one or more samples plus the [cin][tap][cout] weights of a cout tile in shared memory; each
// thread owns 4 output channels x P consecutive pixels of one row.
// This is the exact-fp32 path ("strict" numerics, also the validation reference for the
// tensor-core path).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ktimer.h"
#include "small_tower.h"
#include "launch.h"
#include "conv_tc.h"
#include "pipeline.h"

namespace mz {

// Board layouts of the tensor-core towers (conv_tc.cu / conv_x3.cu): kLayoutF16 = one fp16 plane of 4096 halves,
// kLayoutSplit = two fp16 planes, x_h then x_l with x ~ x_h + x_l / 2^11 (8192 halves = 4096 float slots per board).
enum { kLayoutDense = 0, kLayoutF16 = 1, kLayoutSplit = 2 };
constexpr float kSplitLoScale = 2048.0f, kSplitLoUnscale = 1.0f / 2048.0f;

__device__ __forceinline__ float sat_f16_range(float v) { return fminf(fmaxf(v, -65504.0f), 65504.0f); }
__device__ __forceinline__ void split_f32(float v, __half* hi, __half* lo) {
    const __half h = __float2half_rn(sat_f16_range(v));
    *hi = h;
    *lo = __float2half_rn(sat_f16_range((v - __half2float(h)) * kSplitLoScale));
}

// ------------------------------------------------------------------------------------------
// conv3x3 (+folded BN bias, +residual, +ReLU), stride 1 or 2, pad 1
// ------------------------------------------------------------------------------------------
struct ConvArgs {
    const float* in;          // [n, Cin_real, Hin, Win]   (or gathered from the hidden pool)
    float* out;               // [n, Cout, Ho, Wo]
    const float* residual;    // same shape as out or nullptr
    const float* w;           // [Cin][9][Cout]
    const float* bias;        // [Cout] or nullptr
    const int32_t* gather_parent;   // pool mode: sample g reads in + (g*pool_stride + gather_parent[g]) * sample_elems
    const int32_t* action;    // extra constant input plane action/A as channel Cin-1 (dynamics), or nullptr
    int pool_stride;
    int n, Cin, Cout, Hin, Win, Ho, Wo, stride, relu, A;
    int boards_per_cta, cin_chunk;
    int band_rows;            // output rows per CTA (blockIdx.z picks the band); = Ho unless the image is too large for one CTA
    int out_p64c4;            // kLayoutF16 / kLayoutSplit: write the tensor-core board layout (P64S) instead of NCHW
};

template <int P, int STRIDE, int MAX_ITEMS>
__global__ void __launch_bounds__(256) conv3x3_kernel(const __grid_constant__ ConvArgs a) {
    extern __shared__ __align__(16) float smem[];
    constexpr int IN_SPAN = (P - 1) * STRIDE + 3;              // input columns feeding P outputs
    // this CTA's band of output rows [r0, r0 + nbr) and the input rows it needs: [r0*STRIDE - 1, (r0+nbr-1)*STRIDE + 1]
    const int r0 = blockIdx.z * a.band_rows;
    const int nbr = min(a.band_rows, a.Ho - r0);
    const int y_in0 = r0 * STRIDE - 1;
    const int Hp = (a.band_rows - 1) * STRIDE + 3, Wp = a.Win + 2;   // staged (padded) plane of the band
    const int plane = Hp * Wp;
    const int ct = a.Cout < 64 ? a.Cout : 64;                   // cout tile of this CTA
    const int cout0 = blockIdx.y * ct;
    const int cgs = ct / 4;
    const int segs = a.Wo / P;
    const int items_per_board = cgs * nbr * segs;
    const int b0 = blockIdx.x * a.boards_per_cta;
    const int nb = min(a.boards_per_cta, a.n - b0);
    float* s_w = smem;                                          // [cin_chunk][9][ct]
    float* s_in = smem + a.cin_chunk * 9 * ct;                  // [boards][cin_chunk][Hp][Wp]
    const int cin_real = a.action ? a.Cin - 1 : a.Cin;
    const size_t sample_elems = (size_t)cin_real * a.Hin * a.Win;

    const int total_items = nb * items_per_board;
    // each thread may own several items (big images): MAX_ITEMS accumulator tiles
    float acc[MAX_ITEMS][4][P];
#pragma unroll
    for (int it = 0; it < MAX_ITEMS; ++it)
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int p = 0; p < P; ++p) acc[it][c][p] = 0.0f;

    for (int c0 = 0; c0 < a.Cin; c0 += a.cin_chunk) {
        const int cc = min(a.cin_chunk, a.Cin - c0);
        __syncthreads();
        // ---- stage weights of this cin chunk / cout tile
        for (int i = threadIdx.x; i < cc * 9 * ct; i += blockDim.x) {
            const int co = i % ct, r = i / ct;                  // r = ci*9 + tap
            s_w[i] = a.w[((size_t)(c0 * 9 + r)) * a.Cout + cout0 + co];
        }
        // ---- stage padded input planes
        for (int i = threadIdx.x; i < nb * cc * plane; i += blockDim.x) {
            const int x = i % Wp, y = (i / Wp) % Hp, ci = (i / plane) % cc, b = i / (plane * cc);
            const int g = b0 + b, cg = c0 + ci;
            float v = 0.0f;
            const int yi = y_in0 + y;                           // input row of staged row y
            if (x >= 1 && x <= a.Win && yi >= 0 && yi < a.Hin) {
                if (a.action && cg == a.Cin - 1) {
                    v = __fdiv_rn((float)a.action[g], (float)a.A);       // action / |A| plane
                } else {
                    const float* src = a.gather_parent
                        ? a.in + ((size_t)g * a.pool_stride + a.gather_parent[g]) * sample_elems
                        : a.in + (size_t)g * sample_elems;
                    v = src[((size_t)cg * a.Hin + yi) * a.Win + (x - 1)];
                }
            }
            s_in[i] = v;
        }
        __syncthreads();
        // ---- accumulate
#pragma unroll
        for (int it = 0; it < MAX_ITEMS; ++it) {
            const int item = threadIdx.x + it * blockDim.x;
            if (item >= total_items) break;
            const int cgi = item % cgs;
            const int rest = item / cgs;
            const int seg = rest % segs, y = (rest / segs) % nbr, b = rest / (segs * nbr);
            const float* ib = s_in + (size_t)b * cc * plane + (y * STRIDE) * Wp + seg * P * STRIDE;
            const float* wb = s_w + cgi * 4;
            for (int ci = 0; ci < cc; ++ci) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
                    float v[IN_SPAN];
#pragma unroll
                    for (int j = 0; j < IN_SPAN; ++j) v[j] = ib[ci * plane + dy * Wp + j];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wb + (ci * 9 + dy * 3 + dx) * ct);
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const float xv = v[p * STRIDE + dx];
                            acc[it][0][p] = fmaf(xv, w4.x, acc[it][0][p]);
                            acc[it][1][p] = fmaf(xv, w4.y, acc[it][1][p]);
                            acc[it][2][p] = fmaf(xv, w4.z, acc[it][2][p]);
                            acc[it][3][p] = fmaf(xv, w4.w, acc[it][3][p]);
                        }
                    }
                }
            }
        }
    }
    // ---- epilogue
#pragma unroll
    for (int it = 0; it < MAX_ITEMS; ++it) {
        const int item = threadIdx.x + it * blockDim.x;
        if (item >= total_items) break;
        const int cgi = item % cgs;
        const int rest = item / cgs;
        const int seg = rest % segs, y = r0 + (rest / segs) % nbr, b = rest / (segs * nbr);
        const int g = b0 + b;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int co = cout0 + cgi * 4 + c;
            const float bias = a.bias ? a.bias[co] : 0.0f;
            const size_t o = (((size_t)g * a.Cout + co) * a.Ho + y) * a.Wo + seg * P;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float r = acc[it][c][p] + bias;
                if (a.residual) r += a.residual[o + p];
                if (a.relu) r = fmaxf(r, 0.0f);
                if (a.out_p64c4) {
                    const int pos = (y + 1) * 8 + seg * P + p;
                    const int e = pos * 64 + (((co >> 3) ^ (pos & 7)) << 3) + (co & 7);
                    if (a.out_p64c4 == kLayoutSplit) {
                        __half* base = reinterpret_cast<__half*>(a.out) + (size_t)g * 8192;
                        split_f32(r, base + e, base + 4096 + e);
                    } else {
                        reinterpret_cast<__half*>(a.out)[(size_t)g * 4096 + e] = __float2half_rn(r);
                    }
                }
                else
                    a.out[o + p] = r;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// AvgPool2d(kernel 3, stride 2, padding 1), count_include_pad (always / 9)   models.py:258,262
// ------------------------------------------------------------------------------------------
__global__ void avgpool3x3s2_kernel(const float* in, float* out, int planes, int H, int W, int Ho, int Wo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)planes * Ho * Wo) return;
    const int x = i % Wo, y = (i / Wo) % Ho;
    const size_t pl = i / ((size_t)Wo * Ho);
    const float* p = in + pl * H * W;
    float s = 0.0f;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = 2 * y + dy, xx = 2 * x + dx;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += p[yy * W + xx];
        }
    out[i] = __fdiv_rn(s, 9.0f);
}

// ------------------------------------------------------------------------------------------
// Heads: conv1x1(+bias) -> flatten (c,h,w) -> MLP -> logits (-> support_to_scalar), plus the
// per-(sample, channel) min-max rescale of the state (models.py:530-553).  One CTA per sample.
// ------------------------------------------------------------------------------------------
struct HeadDesc {
    int rc;                    // reduced channels
    int w1_off, b1_off;        // conv1x1 weight [rc][C], bias [rc]
    MlpDesc mlp;               // transposed layers in the same blob
    int n_out;                 // logits
};

struct HeadsArgs {
    const float* x;            // [n, C, HW] input state (raw trunk output)
    const float* blob;
    int n, C, HW, S;
    int n_heads;
    HeadDesc head[2];
    float* logits[2];          // [n, n_out] or nullptr
    float* scalar[2];          // [n] support_to_scalar or nullptr
    // optional rescale of x into the hidden pool / a plain buffer
    float* rescaled;           // [n, C*HW] or nullptr
    float* pool_hidden;        // pool mode target
    int pool_stride, out_slot;
    int smem_floats;
    int p64c4, W;              // kLayoutF16 / kLayoutSplit: input (and pool target) use the tensor-core board layout
    float* state_p64c4;        // [n, 4096 fp16] rescaled state in P64C8 (input of the prediction tower), or nullptr
    int w_lo, w_floats;        // slice of the head blob this launch needs (staged in shared memory)
    int warp_floats;           // per-warp scratch: x tile + two activation vectors
};

// offset (in fp16 elements) of (channel c, dense position p) inside one P64S state of 4096 halves: position-major
// rows of 64 channels, the 8-channel chunks of a row XOR-ed with (padded position % 8) (conv_tc.cu)
__device__ __forceinline__ int p64c4_index(int c, int p, int W) {
    const int pos = (p / W + 1) * 8 + (p % W);
    return pos * 64 + ((((c >> 3) ^ (pos & 7))) << 3) + (c & 7);
}

// Persistent CTAs (one per SM), 1024 threads = 8 groups of 128: the head weights of this launch are staged
// in shared memory once per CTA, then every GROUP takes one sample at a time (named barriers, groups never
// wait for each other).  x is staged as a [position][channel] tile with 16-byte aligned rows (row stride C+4:
// conflict-free for 128-bit row reads and for per-channel column scans).  Everything that touches global
// memory or the weights moves 16 bytes per instruction: the P64S state is read and written as whole 8-channel
// chunks, conv1x1 reads x rows and weight rows as float4, the FC layers read packed [in/4][out][4] weights and
// float4 activations; index arithmetic with runtime divisors happens once per chunk, not per element.
// The accumulation order of every dot product is ascending input index (as torch's reference loops are
// compared with a tolerance anyway, this only keeps results independent of the vector width).
// GROUP = 128 threads per sample for wide states (Connect4: 64 x 42), GROUP = 32 (one warp per sample, __syncwarp
// instead of named barriers, 4x the samples in flight) when a sample is only a few hundred values (TicTacToe 16 x 9,
// Breakout's 16 x 36 hidden board).
constexpr int kHeadThreads = 1024;

template <int GROUP>
__device__ __forceinline__ void group_bar(int group) {
    if constexpr (GROUP == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(GROUP) : "memory");
}

template <int GROUP>
__global__ void __launch_bounds__(kHeadThreads) heads_kernel(const __grid_constant__ HeadsArgs a) {
    constexpr int kHeadGroup = GROUP;
    extern __shared__ __align__(16) float sm[];
    const int C = a.C, HW = a.HW, CP = C + 4;
    const int group = threadIdx.x / kHeadGroup, t = threadIdx.x % kHeadGroup, ngroups = blockDim.x / kHeadGroup;
    float* s_w = sm;                                                 // head blob slice [w_lo, w_lo + w_floats)
    float* s_x = s_w + a.w_floats + (size_t)group * a.warp_floats;   // [HW][C+4]
    float* s_lo = s_x + HW * CP;                                     // [C] channel minimum
    float* s_sc = s_lo + C;                                          // [C] channel scale
    float* s_part = s_sc + C;                                        // [2][2][C] partial extrema
    float* s_act = s_part + 4 * C;                                   // per head: ping | pong
    // board layout only (C = 64): padded row of dense position p, looked up instead of two runtime divisions per chunk
    __shared__ unsigned char s_pos[64];
    pdl_launch_dependents();
    {
        const float4* src = reinterpret_cast<const float4*>(a.blob + a.w_lo);
        float4* dst = reinterpret_cast<float4*>(s_w);
        for (int i = threadIdx.x; i < a.w_floats / 4; i += blockDim.x) dst[i] = src[i];
        if (a.p64c4 && threadIdx.x < HW && threadIdx.x < 64)
            s_pos[threadIdx.x] = (unsigned char)((threadIdx.x / a.W + 1) * 8 + (threadIdx.x % a.W));
    }
    __syncthreads();
    pdl_wait();                                                      // the weights are constants; x comes from the previous kernel
    const float* blob = s_w - a.w_lo;                                // blob[off] addresses the staged copy
    constexpr int cj = 8;                                            // 8-channel chunks per position (board layout: C = 64)
    const int c_shift = (C & (C - 1)) == 0 ? 31 - __clz(C) : -1;     // C is a power of two for every bundled network

    for (int g = blockIdx.x * ngroups + group; g < a.n; g += gridDim.x * ngroups) {
        // ---- stage x[p][c]
        if (a.p64c4) {
            const bool split = a.p64c4 == kLayoutSplit;
            const uint4* x8 = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.x) + (size_t)g * (split ? 8192 : 4096));
            for (int i = t; i < cj * HW; i += kHeadGroup) {
                const int j = i % cj, p = i / cj;
                const int pos = s_pos[p];
                const uint4 v = x8[pos * 8 + (j ^ (pos & 7))];
                const __half2* h2 = reinterpret_cast<const __half2*>(&v);
                float2 f0 = __half22float2(h2[0]), f1 = __half22float2(h2[1]);
                float2 f2 = __half22float2(h2[2]), f3 = __half22float2(h2[3]);
                if (split) {                                   // x = x_h + x_l / 2^11 (second plane)
                    const uint4 w = x8[512 + pos * 8 + (j ^ (pos & 7))];
                    const __half2* l2 = reinterpret_cast<const __half2*>(&w);
                    const float2 g0 = __half22float2(l2[0]), g1 = __half22float2(l2[1]);
                    const float2 g2 = __half22float2(l2[2]), g3 = __half22float2(l2[3]);
                    f0.x = fmaf(g0.x, kSplitLoUnscale, f0.x); f0.y = fmaf(g0.y, kSplitLoUnscale, f0.y);
                    f1.x = fmaf(g1.x, kSplitLoUnscale, f1.x); f1.y = fmaf(g1.y, kSplitLoUnscale, f1.y);
                    f2.x = fmaf(g2.x, kSplitLoUnscale, f2.x); f2.y = fmaf(g2.y, kSplitLoUnscale, f2.y);
                    f3.x = fmaf(g3.x, kSplitLoUnscale, f3.x); f3.y = fmaf(g3.y, kSplitLoUnscale, f3.y);
                }
                float4* d = reinterpret_cast<float4*>(s_x + p * CP + 8 * j);
                d[0] = make_float4(f0.x, f0.y, f1.x, f1.y);
                d[1] = make_float4(f2.x, f2.y, f3.x, f3.y);
            }
        } else {
            const float* x = a.x + (size_t)g * C * HW;
            for (int i = t; i < C * HW; i += kHeadGroup) s_x[(i % HW) * CP + i / HW] = x[i];
        }
        group_bar<GROUP>(group);

        if (a.rescaled || a.pool_hidden || a.state_p64c4) {
            // (x - min) / scale per channel over the positions (models.py:530-553).
            // Phase A: channel extrema (two threads per channel when the group is wide enough).
            const int parts = (2 * C <= kHeadGroup) ? 2 : 1;
            for (int i = t; i < parts * C; i += kHeadGroup) {
                const int c = c_shift >= 0 ? (i & (C - 1)) : i % C, part = c_shift >= 0 ? (i >> c_shift) : i / C;
                const int p0 = (part * HW) / parts, p1 = ((part + 1) * HW) / parts;
                float lo = INFINITY, hi = -INFINITY;
                for (int p = p0; p < p1; ++p) { const float v = s_x[p * CP + c]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
                s_part[(part * 2) * C + c] = lo;
                s_part[(part * 2 + 1) * C + c] = hi;
            }
            group_bar<GROUP>(group);
            for (int c = t; c < C; c += kHeadGroup) {
                float lo = s_part[c], hi = s_part[C + c];
                if (parts == 2) { lo = fminf(lo, s_part[2 * C + c]); hi = fmaxf(hi, s_part[3 * C + c]); }
                float sc = __fsub_rn(hi, lo);
                if (sc < 1e-5f) sc = __fadd_rn(sc, 1e-5f);
                s_lo[c] = lo; s_sc[c] = sc;
            }
            group_bar<GROUP>(group);
            // Phase B: normalise and store
            if (a.p64c4) {
                for (int i = t; i < cj * HW; i += kHeadGroup) {
                    const int j = i % cj, p = i / cj;
                    const int pos = s_pos[p];
                    const float4* xr = reinterpret_cast<const float4*>(s_x + p * CP + 8 * j);
                    const float4* lr = reinterpret_cast<const float4*>(s_lo + 8 * j);
                    const float4* sr = reinterpret_cast<const float4*>(s_sc + 8 * j);
                    float v[8];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 x4 = xr[q], l4 = lr[q], c4 = sr[q];
                        v[4 * q + 0] = div_pos_or_zero(__fsub_rn(x4.x, l4.x), c4.x);
                        v[4 * q + 1] = div_pos_or_zero(__fsub_rn(x4.y, l4.y), c4.y);
                        v[4 * q + 2] = div_pos_or_zero(__fsub_rn(x4.z, l4.z), c4.z);
                        v[4 * q + 3] = div_pos_or_zero(__fsub_rn(x4.w, l4.w), c4.w);
                    }
                    if (a.rescaled) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) a.rescaled[(size_t)g * C * HW + (8 * j + e) * HW + p] = v[e];
                    }
                    uint4 packed, packed_lo;                        // 16-bit operands of the tensor-core convs
                    __half2* h2 = reinterpret_cast<__half2*>(&packed);
                    __half2* l2 = reinterpret_cast<__half2*>(&packed_lo);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {                   // values are in [0, 1]: no range concerns
                        h2[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                        const float2 back = __half22float2(h2[e]);
                        l2[e] = __floats2half2_rn((v[2 * e] - back.x) * kSplitLoScale, (v[2 * e + 1] - back.y) * kSplitLoScale);
                    }
                    const bool split = a.p64c4 == kLayoutSplit;
                    const int off8 = pos * 8 + (j ^ (pos & 7));     // in 16-byte units inside the state's first plane
                    const size_t board16 = split ? 1024 : 512;      // 16-byte units per stored state
                    if (a.pool_hidden) {
                        uint4* dst = reinterpret_cast<uint4*>(a.pool_hidden) + ((size_t)g * a.pool_stride + a.out_slot) * board16;
                        dst[off8] = packed;
                        if (split) dst[512 + off8] = packed_lo;
                    }
                    if (a.state_p64c4) {
                        uint4* dst = reinterpret_cast<uint4*>(a.state_p64c4) + (size_t)g * board16;
                        dst[off8] = packed;
                        if (split) dst[512 + off8] = packed_lo;
                    }
                }
            } else {
                for (int i = t; i < C * HW; i += kHeadGroup) {
                    const int c = i / HW, p = i % HW;
                    const float v = div_pos_or_zero(__fsub_rn(s_x[p * CP + c], s_lo[c]), s_sc[c]);
                    if (a.rescaled) a.rescaled[(size_t)g * C * HW + i] = v;
                    if (a.pool_hidden) a.pool_hidden[((size_t)g * a.pool_stride + a.out_slot) * C * HW + i] = v;
                }
            }
        }

        if (a.n_heads > 0) {
            // the heads of this launch side by side: head h owns threads [h*span, (h+1)*span)
            const int span = kHeadGroup / a.n_heads;
            const int h = t / span, u = t % span;
            const HeadDesc& d = a.head[h];
            float* cur = s_act + (size_t)h * 2 * a.smem_floats;
            float* nxt = cur + a.smem_floats;
            // conv1x1: r[c][p] = b[c] + sum_k W[c][k] x[p][k]; one thread per position, 4 channels at a time
            for (int p = u; p < HW; p += span) {
                const float4* xr = reinterpret_cast<const float4*>(s_x + p * CP);
                for (int c0 = 0; c0 < d.rc; c0 += 4) {
                    const int nc = min(4, d.rc - c0);
                    float acc[4];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) acc[cc] = cc < nc ? blob[d.b1_off + c0 + cc] : 0.0f;
                    const float4* w0 = reinterpret_cast<const float4*>(blob + d.w1_off + (size_t)c0 * C);
                    for (int k4 = 0; k4 < C / 4; ++k4) {
                        const float4 x4 = xr[k4];
#pragma unroll
                        for (int cc = 0; cc < 4; ++cc) {
                            if (cc < nc) {
                                const float4 w4 = w0[cc * (C / 4) + k4];
                                acc[cc] = fmaf(w4.x, x4.x, acc[cc]);
                                acc[cc] = fmaf(w4.y, x4.y, acc[cc]);
                                acc[cc] = fmaf(w4.z, x4.z, acc[cc]);
                                acc[cc] = fmaf(w4.w, x4.w, acc[cc]);
                            }
                        }
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                        if (cc < nc) cur[(c0 + cc) * HW + p] = acc[cc];      // flatten order (c, h, w) = NCHW view(-1, ...)
                }
            }
            if (u < 4) { const int i = d.rc * HW + u; if (i < ((d.rc * HW + 3) & ~3)) cur[i] = 0.0f; }   // zero the padding
            group_bar<GROUP>(group);
            const int max_layers = max(a.head[0].mlp.n, a.head[a.n_heads - 1].mlp.n);
            for (int l = 0; l < max_layers; ++l) {
                if (l < d.mlp.n) {
                    const int in4 = (d.mlp.in[l] + 3) >> 2, out = d.mlp.out[l];
                    const float4* W4 = reinterpret_cast<const float4*>(blob + d.mlp.w_off[l]);      // [in/4][out][4]
                    const float4* a4 = reinterpret_cast<const float4*>(cur);
                    const float* b = blob + d.mlp.b_off[l];
                    const bool last = l == d.mlp.n - 1;
                    for (int o = u; o < ((out + 3) & ~3); o += span) {
                        if (o < out) {
                            float acc = b[o];
#pragma unroll 4
                            for (int i = 0; i < in4; ++i) {
                                const float4 x4 = a4[i], w4 = W4[(size_t)i * out + o];
                                acc = fmaf(x4.x, w4.x, acc);
                                acc = fmaf(x4.y, w4.y, acc);
                                acc = fmaf(x4.z, w4.z, acc);
                                acc = fmaf(x4.w, w4.w, acc);
                            }
                            nxt[o] = last ? acc : elu1(acc);
                        } else {
                            nxt[o] = 0.0f;                          // padding read by the next layer's float4 loads
                        }
                    }
                    float* tmp = cur; cur = nxt; nxt = tmp;
                }
                group_bar<GROUP>(group);
            }
            if (a.logits[h])
                for (int o = u; o < d.n_out; o += span) a.logits[h][(size_t)g * d.n_out + o] = cur[o];
            if (a.scalar[h]) {
                if (span >= 32) {                          // span is a multiple of 32: the head's first warp
                    if (u < 32) {
                        const float v = support_to_scalar_group<32>(cur, a.S);
                        if (u == 0) a.scalar[h][g] = v;
                    }
                } else {                                   // two heads share a warp: 16 lanes each
                    const float v = support_to_scalar_group<16>(cur, a.S);
                    if (u == 0) a.scalar[h][g] = v;
                }
            }
        }
        group_bar<GROUP>(group);
    }
}

// ------------------------------------------------------------------------------------------
// Generic heads for nets whose head weights do not fit in shared memory (games/atari.py: 256 reduced channels x 36
// positions -> FC 9216 -> 256 -> 256 -> 601, 9.4 MB for the first FC layer alone): the same operations as heads_kernel,
// one plain kernel per stage, every dot product accumulated in the same ascending order (so the two routes agree bit
// for bit where both apply).  Throughput is not the point here - availability of the large configuration is.
// ------------------------------------------------------------------------------------------
__global__ void big_rescale_kernel(const float* x, int n, int C, int HW, float* rescaled, float* pool_hidden, int pool_stride, int out_slot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;           // (sample, channel)
    if (i >= n * C) return;
    const int g = i / C;
    const float* src = x + (size_t)i * HW;
    float lo = INFINITY, hi = -INFINITY;
    for (int p = 0; p < HW; ++p) { lo = fminf(lo, src[p]); hi = fmaxf(hi, src[p]); }
    float sc = __fsub_rn(hi, lo);
    if (sc < 1e-5f) sc = __fadd_rn(sc, 1e-5f);
    for (int p = 0; p < HW; ++p) {
        const float v = div_pos_or_zero(__fsub_rn(src[p], lo), sc);
        if (rescaled) rescaled[(size_t)i * HW + p] = v;
        if (pool_hidden) pool_hidden[((size_t)g * pool_stride + out_slot) * C * HW + (size_t)(i % C) * HW + p] = v;
    }
}
// r[g][c][p] = b[c] + sum_k W[c][k] x[g][k][p]
__global__ void big_conv1x1_kernel(const float* x, const float* w, const float* b, int n, int C, int rc, int HW, float* out, int out_stride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * rc * HW) return;
    const int p = i % HW, c = (i / HW) % rc;
    const size_t g = i / ((size_t)HW * rc);
    const float* xs = x + g * C * HW + p;
    const float* ws = w + (size_t)c * C;
    float acc = b[c];
    for (int k = 0; k < C; ++k) acc = fmaf(ws[k], xs[(size_t)k * HW], acc);
    out[g * out_stride + (size_t)c * HW + p] = acc;
}
// y[g][o] = act(b[o] + sum_i x[g][i] W[i][o]), W packed [in/4][out][4]; x rows are `in_stride` apart and zero padded to 4
__global__ void big_fc_kernel(const float* x, const float* W, const float* b, int n, int in, int out, int in_stride, int out_stride,
                              int elu, float* y) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int out4 = (out + 3) & ~3;
    if (i >= (size_t)n * out4) return;
    const int o = i % out4;
    const size_t g = i / out4;
    float r = 0.0f;
    if (o < out) {
        const float4* x4 = reinterpret_cast<const float4*>(x + g * in_stride);
        const float4* W4 = reinterpret_cast<const float4*>(W);
        float acc = b[o];
        const int in4 = (in + 3) >> 2;
        for (int k = 0; k < in4; ++k) {
            const float4 xv = x4[k], wv = W4[(size_t)k * out + o];
            acc = fmaf(xv.x, wv.x, acc);
            acc = fmaf(xv.y, wv.y, acc);
            acc = fmaf(xv.z, wv.z, acc);
            acc = fmaf(xv.w, wv.w, acc);
        }
        r = elu ? elu1(acc) : acc;
    }
    y[g * out_stride + o] = r;                                     // the padding entries read by the next layer are zero
}
__global__ void big_scalar_kernel(const float* logits, int n, int stride, int n_out, int S, float* logits_out, float* scalar) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n) return;
    const float* l = logits + (size_t)g * stride;
    if (logits_out) for (int o = lane; o < n_out; o += 32) logits_out[(size_t)g * n_out + o] = l[o];
    if (scalar) {
        const float v = support_to_scalar_group<32>(l, S);
        if (lane == 0) scalar[g] = v;
    }
}

__global__ void fill_root_reward_logits_kernel(float* out, int n, int F, int S) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * F) out[i] = (i % F == S) ? 0.0f : -INFINITY;
}
__global__ void fill_root_reward_kernel(float* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = inverse_value_transform(0.0f);
}
__global__ void copy_from_pool_kernel(const float* pool, float* out, int n, int pool_stride, int slot, int elems) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * elems) return;
    const size_t g = i / elems, e = i % elems;
    out[i] = pool[(g * pool_stride + slot) * elems + e];
}

// dense fp32 NCHW [count][C][H*W]  <->  fp16 P64C8 [count][C/8][64][8]
__global__ void nchw_to_p64c4_kernel(const float* in, float* out, int count, int C, int H, int W, int split) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int HW = H * W;
    if (i >= (size_t)count * C * HW) return;
    const size_t g = i / ((size_t)C * HW);
    const int c = (i / HW) % C, p = i % HW;
    const int e = p64c4_index(c, p, W);
    if (split) {
        __half* base = reinterpret_cast<__half*>(out) + g * 8192;
        split_f32(in[i], base + e, base + 4096 + e);
    } else {
        reinterpret_cast<__half*>(out)[g * 4096 + e] = __float2half_rn(in[i]);
    }
}
__global__ void p64c4_to_nchw_kernel(const float* in, float* out, int count, int C, int H, int W, int split) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int HW = H * W;
    if (i >= (size_t)count * C * HW) return;
    const size_t g = i / ((size_t)C * HW);
    const int c = (i / HW) % C, p = i % HW;
    const int e = p64c4_index(c, p, W);
    if (split) {
        const __half* base = reinterpret_cast<const __half*>(in) + g * 8192;
        out[i] = fmaf(__half2float(base[4096 + e]), kSplitLoUnscale, __half2float(base[e]));
    } else {
        out[i] = __half2float(reinterpret_cast<const __half*>(in)[g * 4096 + e]);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct ConvLayer {
    int cin, cout, stride;
    size_t w_off;             // into the conv blob: [cin][9][cout]
    long b_off;               // folded BN bias [cout], -1 = none
    long tc_off;              // tensor-core image [9][C/4][C][4] (tf32), -1 = none
    long tc_table_off;        // dynamics first conv: action-plane table [64][C], -1 = none
    long tc_scale_off;        // x3 mode: [C] per-output-channel power of two undoing the weight prescale, -1 = none
};

struct ResNetDevice {
    MzNetDesc net;
    int max_batch, sm_count;
    int C, hh, hw;            // hidden state geometry
    // layers in execution order
    std::vector<ConvLayer> rep_down;   // conv1, rb1 x2 (2 convs each), conv2, rb2 x3, rb3 x3   (downsample only)
    std::vector<ConvLayer> rep_trunk;  // [stem conv if no downsample] + blocks x 2
    std::vector<ConvLayer> dyn;        // conv + blocks x 2
    std::vector<ConvLayer> pred;       // blocks x 2
    HeadDesc reward_head, value_head, policy_head;
    float* d_conv = nullptr;           // conv weights + biases
    float* d_head = nullptr;           // head blob
    float* ws[3] = {nullptr, nullptr, nullptr};
    size_t ws_elems = 0;
    float* scratch_hidden = nullptr;   // [B, C*hh*hw] rescaled state when no pool is given (dense NCHW)
    float* scratch_state = nullptr;    // [B, 4096 fp16] same state in P64C8 (tensor-core path)
    bool loaded = false;
    bool use_tc = false;               // residual towers on tcgen05 (conv_tc.cu / conv_x3.cu)
    bool split = false;                // x3 mode: split operands, fp32-grade accuracy (conv_x3.cu); false = plain fp16 operands
    bool tc_capable = false;           // the shape allows the tensor-core towers at all
    float* big_scratch = nullptr;      // activations of the generic heads route (heads_big)
    size_t big_elems = 0;
    int* d_sat = nullptr;              // x3 mode: number of epilogue threads that stored an activation beyond the fp16 range
    int fell_back = 0;                 // set when the range guard switched this net to the fp32 CUDA-core towers
    bool fuse_small = true;            // CUDA-core towers as one fused launch where they fit (small_tower.cu); MZ_NO_FUSE=1: per layer
    int state_elems = 0;               // float slots per stored hidden state (dense C*H*W, or 2048 = 4096 fp16 for P64C8)
};

static int conv_out(int h, int stride) { return (h - 1) / stride + 1; }

ResNetDevice* resnet_create(const MzNetDesc& net, int max_batch, int sm_count, std::string* err) {
    ResNetDevice* r = new ResNetDevice();
    r->net = net; r->max_batch = max_batch; r->sm_count = sm_count;
    r->C = net.channels;
    if (net.channels % 4 != 0 || (net.downsample && (net.channels / 2) % 4 != 0)) {
        *err = "channels must be a multiple of 4 (8 with downsample)";
        delete r; return nullptr;
    }
    int H = net.obs_h, W = net.obs_w;
    size_t max_elems = (size_t)net.obs_c * H * W;
    if (net.downsample) {
        int h1 = conv_out(H, 2), w1 = conv_out(W, 2);
        int h2 = conv_out(h1, 2), w2 = conv_out(w1, 2);
        int h3 = conv_out(h2, 2), w3 = conv_out(w2, 2);
        int h4 = conv_out(h3, 2), w4 = conv_out(w3, 2);
        max_elems = std::max(max_elems, (size_t)(net.channels / 2) * h1 * w1);
        max_elems = std::max(max_elems, (size_t)net.channels * h2 * w2);
        r->hh = h4; r->hw = w4;
        if (h4 != (H + 15) / 16 || w4 != (W + 15) / 16) {
            *err = "downsample geometry does not match ceil(H/16) x ceil(W/16) (models.py:456-484)";
            delete r; return nullptr;
        }
    } else {
        r->hh = H; r->hw = W;
        max_elems = std::max(max_elems, (size_t)net.channels * H * W);
    }
    max_elems = std::max(max_elems, (size_t)(net.channels + 1) * r->hh * r->hw);
    // MZ_TC_MODE = "off": fp32 CUDA-core towers everywhere; anything else: tcgen05 towers where the shape allows
    // (conv_tc.cu).  MZ_NO_TC=1 is the older spelling of "off".
    const char* no_tc = getenv("MZ_NO_TC");
    const char* tc_mode = getenv("MZ_TC_MODE");
    const bool tc_off = (no_tc && no_tc[0] == '1') || (tc_mode && strcmp(tc_mode, "off") == 0);
    r->tc_capable = !net.downsample && conv_tc_supported(net.channels, r->hh, r->hw);
    r->use_tc = r->tc_capable && !tc_off;
    // default "x3": split 16-bit operands, three partial products, fp32-grade accuracy; "fp16": plain fp16 operands
    // (3x fewer MMAs, ~1e-2 hidden-state error: opt-in)
    r->split = r->use_tc && !(tc_mode && strcmp(tc_mode, "fp16") == 0);
    const char* no_fuse = getenv("MZ_NO_FUSE");
    r->fuse_small = !(no_fuse && no_fuse[0] == '1');
    r->state_elems = r->use_tc ? conv_tc_board_elems(r->split) : r->C * r->hh * r->hw;
    if (r->use_tc) max_elems = std::max(max_elems, (size_t)conv_tc_board_elems(r->split));
    r->ws_elems = max_elems * (size_t)max_batch;
    for (int i = 0; i < 3; ++i) {
        if (cudaMalloc(&r->ws[i], r->ws_elems * 4 + 64) != cudaSuccess) { *err = "workspace allocation failed"; resnet_destroy(r); return nullptr; }
        cudaMemset(r->ws[i], 0, r->ws_elems * 4 + 64);      // P64C4 padding positions must read as zero
    }
    if (cudaMalloc(&r->scratch_hidden, (size_t)max_batch * r->C * r->hh * r->hw * 4 + 64) != cudaSuccess ||
        cudaMalloc(&r->scratch_state, (size_t)max_batch * r->state_elems * 4 + 64) != cudaSuccess) {
        *err = "workspace allocation failed"; resnet_destroy(r); return nullptr;
    }
    cudaMemset(r->scratch_state, 0, (size_t)max_batch * r->state_elems * 4 + 64);
    if (cudaMalloc(&r->d_sat, 64) != cudaSuccess) { *err = "workspace allocation failed"; resnet_destroy(r); return nullptr; }
    cudaMemset(r->d_sat, 0, 64);
    return r;
}

void resnet_destroy(ResNetDevice* r) {
    if (!r) return;
    for (int i = 0; i < 3; ++i) if (r->ws[i]) cudaFree(r->ws[i]);
    if (r->scratch_hidden) cudaFree(r->scratch_hidden);
    if (r->scratch_state) cudaFree(r->scratch_state);
    if (r->d_sat) cudaFree(r->d_sat);
    if (r->big_scratch) cudaFree(r->big_scratch);
    if (r->d_conv) cudaFree(r->d_conv);
    if (r->d_head) cudaFree(r->d_head);
    delete r;
}

namespace {
struct Loader {
    const MzTensor* t; int n; std::string* err; bool ok = true;
    const MzTensor* get(const std::string& name, int64_t numel) {
        for (int i = 0; i < n; ++i)
            if (t[i].name && name == t[i].name) {
                if (t[i].numel != numel) { ok = false; *err = "shape mismatch for " + name; return nullptr; }
                return &t[i];
            }
        ok = false; *err = "missing tensor " + name;
        return nullptr;
    }
};

// conv3x3 [cout][cin][3][3] (+ optional BN prefix) -> [cin][9][cout] with the BN scale folded in
uint16_t to_f16(float x) {          // IEEE fp32 -> fp16, round to nearest even, saturating to +-65504
    uint32_t u;
    memcpy(&u, &x, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    u &= 0x7FFFFFFFu;
    if (u >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);               // >= 65520 (or inf/nan): clamp
    if (u < 0x38800000u) {                                                 // subnormal half or zero
        if (u < 0x33000000u) return (uint16_t)sign;
        const int shift = 126 - (int)(u >> 23);                            // 14..24
        uint32_t mant = (u & 0x7FFFFFu) | 0x800000u;
        const uint32_t lsb = 1u << shift, half = lsb >> 1;
        uint32_t q = mant >> shift;
        const uint32_t rem = mant & (lsb - 1);
        if (rem > half || (rem == half && (q & 1))) ++q;
        return (uint16_t)(sign | q);
    }
    uint32_t e = ((u >> 23) - 112) << 10, m = (u >> 13) & 0x3FFu;
    uint32_t h = e | m;
    const uint32_t rem = u & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;                 // may carry into the exponent: still correct
    return (uint16_t)(sign | h);
}

float f16_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1Fu, m = h & 0x3FFu, u;
    if (e == 0) {
        if (m == 0) u = sign;
        else { int sh = 0; while (!(m & 0x400u)) { m <<= 1; ++sh; } m &= 0x3FFu; u = sign | ((uint32_t)(113 - sh) << 23) | (m << 13); }
    } else if (e == 31) u = sign | 0x7F800000u | (m << 13);
    else u = sign | ((e + 112) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}

bool pack_conv(Loader& L, const std::string& conv, const std::string& bn, int cin, int cout, int stride,
               std::vector<float>& blob, std::vector<ConvLayer>& layers, int tc = 0, int H = 0, int W = 0) {
    const MzTensor* w = L.get(conv + ".weight", (int64_t)cout * cin * 9);
    if (!w) return false;
    std::vector<double> scale(cout, 1.0), shift(cout, 0.0);
    if (!bn.empty()) {
        const MzTensor* g = L.get(bn + ".weight", cout);
        const MzTensor* b = L.get(bn + ".bias", cout);
        const MzTensor* m = L.get(bn + ".running_mean", cout);
        const MzTensor* v = L.get(bn + ".running_var", cout);
        if (!g || !b || !m || !v) return false;
        for (int c = 0; c < cout; ++c) {
            scale[c] = (double)g->data[c] / sqrt((double)v->data[c] + 1e-5);
            shift[c] = (double)b->data[c] - (double)m->data[c] * scale[c];
        }
    }
    ConvLayer l;
    l.cin = cin; l.cout = cout; l.stride = stride;
    l.w_off = blob.size();
    blob.resize(blob.size() + (size_t)cin * 9 * cout);
    float* dst = blob.data() + l.w_off;
    for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
            for (int tap = 0; tap < 9; ++tap)
                dst[((size_t)ci * 9 + tap) * cout + co] = (float)((double)w->data[((size_t)co * cin + ci) * 9 + tap] * scale[co]);
    if (!bn.empty()) {
        l.b_off = (long)blob.size();
        for (int c = 0; c < cout; ++c) blob.push_back((float)shift[c]);
    } else {
        l.b_off = -1;
    }
    while (blob.size() % 4) blob.push_back(0.0f);           // keep every layer 16-byte aligned
    l.tc_off = l.tc_table_off = l.tc_scale_off = -1;
    if (tc == kLayoutSplit) {
        // x3 image [tap][128 rows][cin 64]: rows 0..63 = w_h of cout 0..63, rows 64..127 = w_l; w' = w * 2^k with k per
        // output channel such that the row's largest |w'| lies in [1, 2); w_h = fp16(w'), w_l = fp16(w' - w_h).
        // The 16-byte chunks of a row are XOR-ed with row % 8 (UMMA K-major SWIZZLE_128B).
        const int C = cout;
        l.tc_off = (long)blob.size();
        blob.resize(blob.size() + (size_t)9 * 128 * C / 2);
        l.tc_scale_off = (long)blob.size();
        blob.resize(blob.size() + (size_t)C, 1.0f);
        uint16_t* img = reinterpret_cast<uint16_t*>(blob.data() + l.tc_off);
        float* undo = blob.data() + l.tc_scale_off;
        for (int co = 0; co < C; ++co) {
            float mx = 0.0f;
            for (int ci = 0; ci < C; ++ci)
                for (int tap = 0; tap < 9; ++tap)
                    mx = std::max(mx, fabsf((float)((double)w->data[((size_t)co * cin + ci) * 9 + tap] * scale[co])));
            int k = 0;
            if (mx > 0.0f && std::isfinite(mx)) { int e; frexpf(mx, &e); k = 1 - e; }      // mx * 2^k in [1, 2)
            if (k > 100) k = 100;
            if (k < -100) k = -100;
            undo[co] = ldexpf(1.0f, -k);
            for (int tap = 0; tap < 9; ++tap)
                for (int ci = 0; ci < C; ++ci) {
                    const float wf = (float)((double)w->data[((size_t)co * cin + ci) * 9 + tap] * scale[co]);   // the fp32 path's weight
                    const float ws = ldexpf(wf, k);
                    const uint16_t hb = to_f16(ws);
                    const float back = f16_to_float(hb);
                    const uint16_t lb = to_f16(ws - back);
                    const size_t col = (size_t)((((ci >> 3) ^ (co & 7))) << 3) + (ci & 7);
                    img[((size_t)tap * 128 + co) * C + col] = hb;
                    img[((size_t)tap * 128 + 64 + co) * C + col] = lb;
                }
        }
        if (cin == C + 1) {
            l.tc_table_off = (long)blob.size();
            blob.resize(blob.size() + (size_t)64 * C, 0.0f);
            float* tab = blob.data() + l.tc_table_off;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                    for (int co = 0; co < C; ++co) {
                        double acc = 0.0;
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dx = -1; dx <= 1; ++dx)
                                if (y + dy >= 0 && y + dy < H && x + dx >= 0 && x + dx < W)
                                    acc += (double)w->data[((size_t)co * cin + C) * 9 + (dy + 1) * 3 + (dx + 1)] * scale[co];
                        tab[((y + 1) * 8 + x) * C + co] = (float)acc;
                    }
        }
    } else if (tc) {
        // fp16 image [tap][cout][cin] over the first 64 input channels, the 16-byte chunks of a cout row XOR-ed with
        // cout % 8 (UMMA K-major SWIZZLE_128B; two halves per float slot of the blob); an extra (65th) input channel
        // is the constant action plane and becomes a per-position fp32 table (sum of the taps that stay inside the board)
        const int C = cout;
        l.tc_off = (long)blob.size();
        blob.resize(blob.size() + (size_t)9 * C * C / 2);
        uint16_t* img = reinterpret_cast<uint16_t*>(blob.data() + l.tc_off);
        for (int tap = 0; tap < 9; ++tap)
            for (int ci = 0; ci < C; ++ci)
                for (int co = 0; co < C; ++co)
                    img[((size_t)tap * C + co) * C + ((((ci >> 3) ^ (co & 7))) << 3) + (ci & 7)] =
                        to_f16((float)((double)w->data[((size_t)co * cin + ci) * 9 + tap] * scale[co]));
        if (cin == C + 1) {
            l.tc_table_off = (long)blob.size();
            blob.resize(blob.size() + (size_t)64 * C, 0.0f);
            float* tab = blob.data() + l.tc_table_off;
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x)
                    for (int co = 0; co < C; ++co) {
                        double acc = 0.0;
                        for (int dy = -1; dy <= 1; ++dy)
                            for (int dx = -1; dx <= 1; ++dx)
                                if (y + dy >= 0 && y + dy < H && x + dx >= 0 && x + dx < W)
                                    acc += (double)w->data[((size_t)co * cin + C) * 9 + (dy + 1) * 3 + (dx + 1)] * scale[co];
                        tab[((y + 1) * 8 + x) * C + co] = (float)acc;
                    }
        }
    }
    layers.push_back(l);
    return true;
}

bool pack_resblock(Loader& L, const std::string& p, int ch, std::vector<float>& blob, std::vector<ConvLayer>& layers,
                   int tc = 0) {
    return pack_conv(L, p + ".conv1", p + ".bn1", ch, ch, 1, blob, layers, tc) &&
           pack_conv(L, p + ".conv2", p + ".bn2", ch, ch, 1, blob, layers, tc);
}

bool pack_head(Loader& L, const std::string& conv, const std::string& fc, int C, int rc, int HW, const int32_t* hidden,
               int n_hidden, int n_out, std::vector<float>& blob, HeadDesc& d) {
    const MzTensor* w = L.get(conv + ".weight", (int64_t)rc * C);
    const MzTensor* b = L.get(conv + ".bias", rc);
    if (!w || !b) return false;
    d.rc = rc; d.n_out = n_out;
    while (blob.size() % 4) blob.push_back(0.0f);          // every head starts 16-byte aligned (float4 staging)
    d.w1_off = (int)blob.size(); blob.insert(blob.end(), w->data, w->data + (size_t)rc * C);
    d.b1_off = (int)blob.size(); blob.insert(blob.end(), b->data, b->data + rc);
    std::vector<int> sz;
    sz.push_back(rc * HW);
    for (int i = 0; i < n_hidden; ++i) sz.push_back(hidden[i]);
    sz.push_back(n_out);
    d.mlp.n = (int)sz.size() - 1;
    for (int l = 0; l < d.mlp.n; ++l) {
        const int in = sz[l], out = sz[l + 1];
        const MzTensor* lw = L.get(fc + "." + std::to_string(2 * l) + ".weight", (int64_t)in * out);
        const MzTensor* lb = L.get(fc + "." + std::to_string(2 * l) + ".bias", out);
        if (!lw || !lb) return false;
        d.mlp.in[l] = in; d.mlp.out[l] = out;
        while (blob.size() % 4) blob.push_back(0.0f);
        d.mlp.w_off[l] = (int)blob.size();
        const int in4 = (in + 3) / 4;
        blob.resize(blob.size() + (size_t)in4 * out * 4, 0.0f);        // packed [in/4][out][4], zero rows pad `in`
        float* dst = blob.data() + d.mlp.w_off[l];
        for (int o = 0; o < out; ++o)
            for (int i = 0; i < in; ++i) dst[((size_t)(i / 4) * out + o) * 4 + (i % 4)] = lw->data[(size_t)o * in + i];
        d.mlp.b_off[l] = (int)blob.size();
        blob.insert(blob.end(), lb->data, lb->data + out);
    }
    while (blob.size() % 4) blob.push_back(0.0f);
    return true;
}
}  // namespace

int resnet_load_weights(ResNetDevice* r, const MzTensor* tensors, int n, std::string* err) {
    const MzNetDesc& nd = r->net;
    const int C = nd.channels, HW = r->hh * r->hw, F = 2 * nd.support_size + 1;
    Loader L{tensors, n, err};
    std::vector<float> conv, head;
    r->rep_down.clear(); r->rep_trunk.clear(); r->dyn.clear(); r->pred.clear();
    const std::string rp = "representation_network.module";
    bool ok = true;
    if (nd.downsample) {
        const std::string dp = rp + ".downsample_net";
        ok = ok && pack_conv(L, dp + ".conv1", "", nd.obs_c, C / 2, 2, conv, r->rep_down);
        for (int i = 0; ok && i < 2; ++i) ok = pack_resblock(L, dp + ".resblocks1." + std::to_string(i), C / 2, conv, r->rep_down);
        ok = ok && pack_conv(L, dp + ".conv2", "", C / 2, C, 2, conv, r->rep_down);
        for (int i = 0; ok && i < 3; ++i) ok = pack_resblock(L, dp + ".resblocks2." + std::to_string(i), C, conv, r->rep_down);
        for (int i = 0; ok && i < 3; ++i) ok = pack_resblock(L, dp + ".resblocks3." + std::to_string(i), C, conv, r->rep_down);
    } else {
        ok = ok && pack_conv(L, rp + ".conv", rp + ".bn", nd.obs_c, C, 1, conv, r->rep_trunk);
    }
    // the tensor-core images are packed whenever the shape allows them (both the fp16 and the x3 image are cheap), so the
    // range guard can switch paths without reloading; which one is used is decided per launch
    const int tc = !r->tc_capable ? 0 : (r->split ? kLayoutSplit : kLayoutF16);
    for (int i = 0; ok && i < nd.blocks; ++i) ok = pack_resblock(L, rp + ".resblocks." + std::to_string(i), C, conv, r->rep_trunk, tc);
    const std::string dp = "dynamics_network.module";
    ok = ok && pack_conv(L, dp + ".conv", dp + ".bn", C + 1, C, 1, conv, r->dyn, tc, r->hh, r->hw);
    for (int i = 0; ok && i < nd.blocks; ++i) ok = pack_resblock(L, dp + ".resblocks." + std::to_string(i), C, conv, r->dyn, tc);
    const std::string pp = "prediction_network.module";
    for (int i = 0; ok && i < nd.blocks; ++i) ok = pack_resblock(L, pp + ".resblocks." + std::to_string(i), C, conv, r->pred, tc);
    ok = ok && pack_head(L, dp + ".conv1x1_reward", dp + ".fc", C, nd.reduced_reward, HW, nd.res_fc_reward, nd.n_res_fc_reward, F, head, r->reward_head);
    ok = ok && pack_head(L, pp + ".conv1x1_value", pp + ".fc_value", C, nd.reduced_value, HW, nd.res_fc_value, nd.n_res_fc_value, F, head, r->value_head);
    ok = ok && pack_head(L, pp + ".conv1x1_policy", pp + ".fc_policy", C, nd.reduced_policy, HW, nd.res_fc_policy, nd.n_res_fc_policy, nd.action_space, head, r->policy_head);
    if (!ok || !L.ok) return MZ_EINVAL;
    if (r->d_conv) cudaFree(r->d_conv);
    if (r->d_head) cudaFree(r->d_head);
    r->d_conv = r->d_head = nullptr;
    if (cudaMalloc(&r->d_conv, conv.size() * 4 + 64) != cudaSuccess || cudaMalloc(&r->d_head, head.size() * 4 + 64) != cudaSuccess) {
        *err = "weight allocation failed"; return MZ_ENOMEM;
    }
    cudaMemcpy(r->d_conv, conv.data(), conv.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(r->d_head, head.data(), head.size() * 4, cudaMemcpyHostToDevice);
    r->loaded = true;
    return MZ_OK;
}

namespace {
struct Runner {
    ResNetDevice* r; cudaStream_t stream; int64_t* launches; std::string* err; int n;
    bool fail(const char* what, cudaError_t e) { *err = std::string(what) + ": " + cudaGetErrorString(e); return false; }

    // conv: in -> out. `in` may be gathered from the pool; action adds the constant plane.
    TowerLayer tower_layer(const ConvLayer& l, int in_buf, int out_buf, int res_buf, bool relu) {
        TowerLayer t{};
        t.w = r->d_conv + l.tc_off;
        t.bias = l.b_off >= 0 ? r->d_conv + l.b_off : nullptr;
        t.action_table = l.tc_table_off >= 0 ? r->d_conv + l.tc_table_off : nullptr;
        t.scale = l.tc_scale_off >= 0 ? r->d_conv + l.tc_scale_off : nullptr;
        t.in_buf = in_buf; t.out_buf = out_buf; t.res_buf = res_buf; t.relu = relu ? 1 : 0;
        return t;
    }

    bool launch_tower(TowerArgs& a) {
        a.n = n; a.H = r->hh; a.W = r->hw; a.A = r->net.action_space;
        static const int dbg = getenv("MZ_TC_DEBUG_SKIP") ? atoi(getenv("MZ_TC_DEBUG_SKIP")) : 0;
        a.debug_skip = dbg;
        a.g0 = 0; a.sat_count = r->d_sat;
        kt_begin(KT_TOWER, stream);
        cudaError_t e = r->split ? launch_conv_tower_x3(a, r->sm_count, stream) : launch_conv_tower_tc(a, r->sm_count, stream);
        kt_end(stream);
        if (e != cudaSuccess) return fail("conv_tower_tc launch", e);
        *launches += r->split ? conv_x3_launches(n, r->sm_count) : 1;
        return true;
    }

    // one tensor-core conv (a tower of one layer)
    bool conv_tc(const ConvLayer& l, const float* in, float* out, const float* residual, bool relu,
                 const int32_t* gather_parent = nullptr, int pool_stride = 0, const int32_t* action = nullptr) {
        TowerArgs a{};
        a.n_layers = 1;
        a.buf[0] = const_cast<float*>(in); a.buf[1] = out; a.buf[2] = const_cast<float*>(residual);
        a.layer[0] = tower_layer(l, 0, 1, residual ? 2 : -1, relu);
        if (!action) a.layer[0].action_table = nullptr;
        a.gather_parent = gather_parent; a.pool_stride = pool_stride; a.action = action;
        return launch_tower(a);
    }

    bool blocks_tc(const std::vector<ConvLayer>& layers, size_t first, size_t count, float** cur, float** tmp, float** spare) {
        for (size_t b = 0; b < count; ++b) {
            if (!conv_tc(layers[first + 2 * b], *cur, *tmp, nullptr, true)) return false;
            if (!conv_tc(layers[first + 2 * b + 1], *tmp, *spare, *cur, true)) return false;
            float* t = *cur; *cur = *spare; *spare = t;
        }
        return true;
    }

    // [optional stem conv] + `count` residual blocks in as few persistent launches as possible.
    // ext = tower input (read only, optionally gathered from the hidden pool); ws = three workspaces.
    // Returns the buffer holding the result (one of ws, or ext if there was nothing to do) or nullptr on error.
    // `first` = index of the first layer to run; `ext_reusable`: ext is scratch that may be overwritten once read.
    const float* tower_tc(const std::vector<ConvLayer>& layers, size_t first_layer, bool stem, size_t count, const float* ext,
                          bool ext_reusable, float* const ws[3], const int32_t* gather_parent, int pool_stride,
                          const int32_t* action) {
        if (!r->split && n > conv_tc_max_boards_fused(r->sm_count)) {
            // too many tiles per CTA for the fused mode: one launch per conv
            float* free_ws[3]; int nf = 0;
            for (int i = 0; i < 3; ++i) if (ws[i] != ext) free_ws[nf++] = ws[i];
            if (nf < 3) free_ws[nf++] = const_cast<float*>(ext);      // ext is itself a workspace: reusable as the third
            float *cur = free_ws[0], *tmp = free_ws[1], *spare = free_ws[2];
            size_t first = first_layer;
            const float* x = ext;
            if (stem) { if (!conv_tc(layers[first], ext, cur, nullptr, true, gather_parent, pool_stride, action)) return nullptr; first += 1; x = cur; }
            if (count > 0 && !stem) {
                if (!conv_tc(layers[first], ext, tmp, nullptr, true, gather_parent, pool_stride)) return nullptr;
                if (!conv_tc(layers[first + 1], tmp, spare, ext, true)) return nullptr;      // residual = ext (plain addressing only)
                { float* t = cur; cur = spare; spare = t; }
                first += 2; count -= 1; x = cur;
            }
            if (!blocks_tc(layers, first, count, &cur, &tmp, &spare)) return nullptr;
            return x == ext ? x : cur;
        }
        size_t li = first_layer, blocks_left = count;
        bool stem_left = stem;
        const float* ext_now = ext;                 // buf[0] of the next launch
        const int32_t* gather_now = gather_parent;
        const float* result = ext;
        while (stem_left || blocks_left > 0) {
            TowerArgs a{};
            a.buf[0] = const_cast<float*>(ext_now);
            // the three workspaces, skipping the one that currently holds the input
            int nb = 1;
            for (int i = 0; i < 3; ++i) if (ws[i] != ext_now) a.buf[nb++] = ws[i];
            if (nb < 4) a.buf[nb++] = nullptr;      // (only two spare workspaces when the input is one of ws)
            a.gather_parent = gather_now; a.pool_stride = pool_stride; a.action = action;
            int cur = 0, nl = 0;
            const bool reuse0 = ext_reusable && !gather_now;       // the input buffer is dead after its last reader
            auto pick = [&](int avoid1, int avoid2) {
                for (int i = 1; i < 4; ++i) if (i != avoid1 && i != avoid2 && a.buf[i]) return i;
                if (reuse0 && avoid1 != 0 && avoid2 != 0) return 0;
                return -1;
            };
            if (stem_left) {
                const int o = pick(cur, -1);
                a.layer[nl++] = tower_layer(layers[li++], cur, o, -1, true);
                cur = o; stem_left = false;
            }
            while (blocks_left > 0 && nl + 2 <= kTowerMaxLayers) {
                const int t1 = pick(cur, -1);
                const int t2 = pick(cur, t1);
                if (t1 < 0 || t2 < 0) break;
                a.layer[nl++] = tower_layer(layers[li++], cur, t1, -1, true);
                a.layer[nl++] = tower_layer(layers[li++], t1, t2, cur, true);
                cur = t2; --blocks_left;
            }
            if (nl == 0) { *err = "tower_tc: no workspace left"; return nullptr; }
            for (int i = 0; i < nl; ++i) if (!action) a.layer[i].action_table = nullptr;
            a.n_layers = nl;
            if (!launch_tower(a)) return nullptr;
            result = a.buf[cur];
            ext_now = result; gather_now = nullptr;
        }
        return result;
    }

    bool conv(const ConvLayer& l, const float* in, float* out, const float* residual, bool relu, int Hin, int Win,
              const int32_t* gather_parent = nullptr, int pool_stride = 0, const int32_t* action = nullptr,
              bool out_p64c4 = false) {
        ConvArgs a{};
        a.out_p64c4 = out_p64c4 ? (r->split ? kLayoutSplit : kLayoutF16) : 0;
        a.in = in; a.out = out; a.residual = residual; a.w = r->d_conv + l.w_off;
        a.bias = l.b_off >= 0 ? r->d_conv + l.b_off : nullptr;
        a.gather_parent = gather_parent; a.pool_stride = pool_stride; a.action = action;
        a.n = n; a.Cin = l.cin; a.Cout = l.cout; a.Hin = Hin; a.Win = Win; a.stride = l.stride;
        a.Ho = conv_out(Hin, l.stride); a.Wo = conv_out(Win, l.stride); a.relu = relu; a.A = r->net.action_space;
        int P = 1;
        for (int cand : {8, 7, 6, 4, 3, 2}) if (a.Wo % cand == 0) { P = cand; break; }
        const int ct = l.cout < 64 ? l.cout : 64;
        const int threads = 256;
        // large images (e.g. 128 output channels at 48 x 48, games/atari.py): split the output rows into bands, one CTA each
        int bands = 1;
        while (bands < a.Ho && (ct / 4) * ((a.Ho + bands - 1) / bands) * (a.Wo / P) > threads * 4) ++bands;
        a.band_rows = (a.Ho + bands - 1) / bands;
        bands = (a.Ho + a.band_rows - 1) / a.band_rows;
        const int items_per_board = (ct / 4) * a.band_rows * (a.Wo / P);
        int boards = 1;
        if (items_per_board < threads) boards = threads / items_per_board;
        if (boards > 32) boards = 32;
        if (boards > n) boards = n;
        if (items_per_board * boards > threads * 4) { *err = "conv3x3: image too large for the item budget"; return false; }
        const size_t plane = (size_t)((a.band_rows - 1) * l.stride + 3) * (Win + 2);
        // pick the cin chunk so weights + planes fit comfortably
        const size_t budget = 200 * 1024 / 4;
        int chunk = l.cin;
        while (chunk > 1 && (size_t)chunk * 9 * ct + (size_t)boards * chunk * plane > budget) chunk = (chunk + 1) / 2;
        if ((size_t)chunk * 9 * ct + (size_t)boards * chunk * plane > budget) { *err = "conv3x3: tile does not fit in shared memory"; return false; }
        a.boards_per_cta = boards; a.cin_chunk = chunk;
        const size_t smem = ((size_t)chunk * 9 * ct + (size_t)boards * chunk * plane) * 4;
        dim3 grid((n + boards - 1) / boards, l.cout / ct, bands);
        const bool multi = items_per_board * boards > threads;
#define MZ_CONV(PP, SS)                                                                                         \
        if (P == PP && l.stride == SS) {                                                                        \
            auto kern = multi ? conv3x3_kernel<PP, SS, 4> : conv3x3_kernel<PP, SS, 1>;                          \
            static size_t attr_smem[2] = {0, 0};                                                                \
            if (attr_smem[multi] < smem) {                                                                      \
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                if (e != cudaSuccess) return fail("conv attr", e);                                              \
                attr_smem[multi] = smem;                                                                        \
            }                                                                                                   \
            kt_begin(KT_CONV, stream);                                                                          \
            kern<<<grid, threads, smem, stream>>>(a);                                                           \
            kt_end(stream);                                                                                     \
        }
        MZ_CONV(8, 1) MZ_CONV(7, 1) MZ_CONV(6, 1) MZ_CONV(4, 1) MZ_CONV(3, 1) MZ_CONV(2, 1) MZ_CONV(1, 1)
        MZ_CONV(8, 2) MZ_CONV(6, 2) MZ_CONV(4, 2) MZ_CONV(3, 2) MZ_CONV(2, 2) MZ_CONV(1, 2) MZ_CONV(7, 2)
#undef MZ_CONV
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail("conv3x3 launch", e);
        *launches += 1;
        return true;
    }

    // [optional stem conv] + `count` residual blocks as ONE fused CUDA-core launch (small_tower.cu).
    // Returns 1 = launched, 0 = shape not supported (caller falls back to one launch per conv), -1 = error.
    int small_tower(const std::vector<ConvLayer>& layers, size_t first, bool stem, size_t count, const float* in, float* out,
                    int in_channels, int H, int W, const int32_t* gather_parent = nullptr, int pool_stride = 0,
                    const int32_t* action = nullptr) {
        const bool off = !r->fuse_small;
        const size_t nl = (stem ? 1 : 0) + 2 * count;
        if (off || nl == 0 || nl > (size_t)kSmallTowerMaxLayers) return 0;
        SmallTowerArgs a{};
        a.in = in; a.out = out; a.blob = r->d_conv; a.gather_parent = gather_parent; a.action = action; a.pool_stride = pool_stride;
        a.n = n; a.C = r->C; a.H = H; a.W = W; a.A = r->net.action_space; a.in_channels = in_channels; a.n_layers = (int)nl;
        for (size_t i = 0; i < nl; ++i) {
            const ConvLayer& l = layers[first + i];
            if (l.stride != 1 || l.cout != r->C) return 0;
            SmallTowerLayer& t = a.layer[i];
            t.w_off = (int)l.w_off; t.b_off = (int)l.b_off; t.cin = l.cin; t.relu = 1;
            t.residual = (i >= (stem ? 1u : 0u) && ((i - (stem ? 1 : 0)) & 1)) ? 1 : 0;     // second conv of a block
        }
        if (!small_tower_supported(a)) return 0;
        kt_begin(KT_SMALL, stream);
        cudaError_t e = launch_small_tower(a, r->sm_count, stream);
        kt_end(stream);
        if (e != cudaSuccess) { fail("small_tower launch", e); return -1; }
        *launches += 1;
        return 1;
    }

    // residual tower: layers[2k], layers[2k+1] are one block; x ends up in `*cur`
    bool blocks(const std::vector<ConvLayer>& layers, size_t first, size_t count, float** cur, float** tmp, float** spare, int H, int W) {
        for (size_t b = 0; b < count; ++b) {
            if (!conv(layers[first + 2 * b], *cur, *tmp, nullptr, true, H, W)) return false;
            if (!conv(layers[first + 2 * b + 1], *tmp, *spare, *cur, true, H, W)) return false;
            float* t = *cur; *cur = *spare; *spare = t;
        }
        return true;
    }

    // heads whose weights do not fit in shared memory: one plain kernel per stage (see big_*_kernel above)
    bool heads_big(const float* x, int n_heads, const HeadDesc* const* hs, float* l0, float* l1, float* s0, float* s1,
                   float* rescaled, float* pool_hidden, int pool_stride, int out_slot) {
        const int C = r->C, HW = r->hh * r->hw, S = r->net.support_size;
        kt_begin(KT_HEADS, stream);
        if (rescaled || pool_hidden) {
            big_rescale_kernel<<<(n * C + 127) / 128, 128, 0, stream>>>(x, n, C, HW, rescaled, pool_hidden, pool_stride, out_slot);
            *launches += 1;
        }
        float* logits_out[2] = {l0, l1};
        float* scalar_out[2] = {s0, s1};
        for (int hi = 0; hi < n_heads; ++hi) {
            const HeadDesc& d = *hs[hi];
            int width = ((d.rc * HW + 3) & ~3);
            for (int l = 0; l < d.mlp.n; ++l) width = std::max(width, (d.mlp.out[l] + 3) & ~3);
            const size_t need = (size_t)2 * n * width;
            if (r->big_elems < need) {
                if (r->big_scratch) cudaFree(r->big_scratch);
                r->big_scratch = nullptr; r->big_elems = 0;
                if (cudaMalloc(&r->big_scratch, need * 4 + 64) != cudaSuccess) { *err = "heads: scratch allocation failed"; kt_end(stream); return false; }
                r->big_elems = need;
            }
            float* cur = r->big_scratch;
            float* nxt = r->big_scratch + (size_t)n * width;
            cudaMemsetAsync(cur, 0, (size_t)n * width * 4, stream);           // zero padding behind rc*HW
            const size_t items = (size_t)n * d.rc * HW;
            big_conv1x1_kernel<<<(unsigned)((items + 127) / 128), 128, 0, stream>>>(x, r->d_head + d.w1_off, r->d_head + d.b1_off, n, C, d.rc, HW, cur, width);
            *launches += 1;
            for (int l = 0; l < d.mlp.n; ++l) {
                const int out4 = (d.mlp.out[l] + 3) & ~3;
                big_fc_kernel<<<(unsigned)(((size_t)n * out4 + 127) / 128), 128, 0, stream>>>(
                    cur, r->d_head + d.mlp.w_off[l], r->d_head + d.mlp.b_off[l], n, d.mlp.in[l], d.mlp.out[l], width, width,
                    l == d.mlp.n - 1 ? 0 : 1, nxt);
                *launches += 1;
                float* t = cur; cur = nxt; nxt = t;
            }
            big_scalar_kernel<<<(n * 32 + 127) / 128, 128, 0, stream>>>(cur, n, width, d.n_out, S, logits_out[hi], scalar_out[hi]);
            *launches += 1;
        }
        kt_end(stream);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail("heads (generic route)", e);
        return true;
    }

    bool heads(const float* x, int n_heads, const HeadDesc* h0, const HeadDesc* h1, float* l0, float* l1, float* s0, float* s1,
               float* rescaled, float* pool_hidden, int pool_stride, int out_slot, bool p64c4 = false, float* state_p64c4 = nullptr) {
        HeadsArgs a{};
        a.p64c4 = p64c4 ? (r->split ? kLayoutSplit : kLayoutF16) : 0; a.W = r->hw; a.state_p64c4 = state_p64c4;
        a.x = x; a.blob = r->d_head; a.n = n; a.C = r->C; a.HW = r->hh * r->hw; a.S = r->net.support_size;
        a.n_heads = n_heads;
        int maxw = 32;
        const HeadDesc* hs[2] = {h0, h1};
        for (int i = 0; i < n_heads; ++i) {
            a.head[i] = *hs[i];
            maxw = std::max(maxw, hs[i]->rc * a.HW + 4);
            for (int l = 0; l < hs[i]->mlp.n; ++l) maxw = std::max(maxw, hs[i]->mlp.out[l] + 4);
        }
        a.logits[0] = l0; a.logits[1] = l1; a.scalar[0] = s0; a.scalar[1] = s1;
        a.rescaled = rescaled; a.pool_hidden = pool_hidden; a.pool_stride = pool_stride; a.out_slot = out_slot;
        a.smem_floats = (maxw + 3) & ~3;
        // blob slice covering the heads of this launch
        int lo = 1 << 30, hi = 0;
        for (int i = 0; i < n_heads; ++i) {
            const HeadDesc& d = *hs[i];
            lo = std::min(lo, d.w1_off);
            const int last = d.mlp.n - 1;
            hi = std::max(hi, d.mlp.b_off[last] + d.mlp.out[last]);
        }
        if (n_heads == 0) { lo = 0; hi = 0; }
        a.w_lo = lo; a.w_floats = ((hi - lo) + 3) & ~3;
        a.warp_floats = (a.HW * (a.C + 4) + 6 * a.C + 4 * a.smem_floats + 3) & ~3;   // x tile + channel stats + (ping, pong) per head
        // one warp per sample when a sample is small (and, for small batches, only as many groups per CTA as it takes
        // to give every SM work); 128 threads per sample otherwise
        const bool narrow = a.C * a.HW <= 1024 && n_heads <= 2;
        const int group = narrow ? 32 : 128;
        int groups = kHeadThreads / group;
        // only as many groups per CTA as it takes to give every SM work (1024 Connect4 boards: 147 CTAs x 7 groups, not 128 x 8)
        groups = std::max(1, std::min(groups, (n + r->sm_count - 1) / r->sm_count));
        size_t smem = ((size_t)a.w_floats + (size_t)groups * a.warp_floats) * 4;
        while (groups > 1 && smem > 227 * 1024) { --groups; smem = ((size_t)a.w_floats + (size_t)groups * a.warp_floats) * 4; }
        if (smem > 227 * 1024) {
            if (p64c4) { *err = "heads: weights + tiles exceed shared memory"; return false; }
            return heads_big(x, n_heads, hs, l0, l1, s0, s1, rescaled, pool_hidden, pool_stride, out_slot);
        }
        const int threads = groups * group;
        static size_t attr_smem[2] = {0, 0};
        if (attr_smem[narrow] < smem) {
            cudaError_t e0 = narrow ? cudaFuncSetAttribute(heads_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                    : cudaFuncSetAttribute(heads_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e0 != cudaSuccess) return fail("heads attr", e0);
            attr_smem[narrow] = smem;
        }
        int grid = (n + groups - 1) / groups;
        if (grid > r->sm_count) grid = r->sm_count;
        kt_begin(KT_HEADS, stream);
        cudaError_t e = narrow ? launch_chained(heads_kernel<32>, dim3(grid), dim3(threads), smem, stream, a)
                               : launch_chained(heads_kernel<128>, dim3(grid), dim3(threads), smem, stream, a);
        kt_end(stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return fail("heads launch", e);
        *launches += 1;
        return true;
    }
};
}  // namespace

// Tensor-core variant: every C->C conv of the three towers runs in conv_tc.cu on the P64C4 layout; the
// stem conv (obs -> C) and the heads stay on the CUDA-core kernels above, reading / writing that layout.
static int resnet_inference_tc(ResNetDevice* r, const InferCall& c, cudaStream_t stream, int64_t* launches, std::string* err) {
    const MzNetDesc& nd = r->net;
    const int n = c.n, C = r->C, hh = r->hh, hw = r->hw, F = 2 * nd.support_size + 1;
    Runner R{r, stream, launches, err, n};
    float *cur = r->ws[0], *tmp = r->ws[1], *spare = r->ws[2];
    float* state = r->scratch_state;                   // rescaled state, P64C4, input of the prediction tower
    if (!c.recurrent) {
        if (!R.conv(r->rep_trunk[0], c.in, cur, nullptr, true, nd.obs_h, nd.obs_w, nullptr, 0, nullptr, true)) return MZ_ECUDA;
        const float* x = R.tower_tc(r->rep_trunk, 1, false, nd.blocks, cur, true, r->ws, nullptr, 0, nullptr);
        if (!x) return MZ_ECUDA;
        // note: layers index from 1 in rep_trunk (0 is the stem, run above on the CUDA cores)
        if (!R.heads(x, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, c.hidden, c.pool_hidden, c.pool_stride, c.out_slot,
                     true, state))
            return MZ_ECUDA;
        if (c.reward_logits) {
            fill_root_reward_logits_kernel<<<(n * F + 255) / 256, 256, 0, stream>>>(c.reward_logits, n, F, nd.support_size);
            *launches += 1;
        }
        if (c.reward) {
            fill_root_reward_kernel<<<(n + 255) / 256, 256, 0, stream>>>(c.reward, n);
            *launches += 1;
        }
    } else {
        const float* in = c.pool_hidden;
        if (!c.gather_parent) {
            // plain API call: dense NCHW hidden states -> P64C4
            const size_t total = (size_t)n * C * hh * hw;
            nchw_to_p64c4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(c.in, spare, n, C, hh, hw, r->split ? 1 : 0);
            *launches += 1;
            in = spare;
        }
        const float* x = R.tower_tc(r->dyn, 0, true, nd.blocks, in, in == spare, r->ws, c.gather_parent, c.pool_stride, c.action);
        if (!x) return MZ_ECUDA;
        if (!R.heads(x, 1, &r->reward_head, nullptr, c.reward_logits, nullptr, c.reward, nullptr, c.hidden, c.pool_hidden,
                     c.pool_stride, c.out_slot, true, state))
            return MZ_ECUDA;
    }
    const float* x = R.tower_tc(r->pred, 0, false, nd.blocks, state, true, r->ws, nullptr, 0, nullptr);
    if (!x) return MZ_ECUDA;
    if (!R.heads(x, 2, &r->value_head, &r->policy_head, c.value_logits, c.policy_logits, c.value, nullptr, nullptr, nullptr, 0, 0, true))
        return MZ_ECUDA;
    (void)tmp;
    return MZ_OK;
}

int resnet_state_elems(const ResNetDevice* r) { return r->state_elems; }
const char* resnet_numerics(const ResNetDevice* r) {
    if (!r->use_tc) return r->fell_back ? "f32 nets + f64 tree statistics (tensor-core towers left after an activation exceeded the fp16 range)"
                                        : "f32 nets + f64 tree statistics";
    return r->split ? "f32-grade nets (tensor-core towers on split fp16 operands x = x_h + x_l/2^11, 3 partial products, f32 accumulate; f32 heads) + f64 tree statistics"
                    : "fp16 operands / f32 accumulate (tensor-core towers), f32 heads, f64 tree statistics";
}

// Range guard of the x3 towers: number of epilogue threads that stored an activation beyond the fp16 range since the
// last call (synchronises the stream).  resnet_use_strict switches the handle to the fp32 CUDA-core towers for good.
int resnet_take_saturations(ResNetDevice* r, cudaStream_t stream) {
    if (!r->use_tc || !r->split || !r->d_sat) return 0;
    int count = 0;
    if (cudaMemcpyAsync(&count, r->d_sat, 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(stream) != cudaSuccess) return 0;
    if (count) cudaMemsetAsync(r->d_sat, 0, 4, stream);
    return count;
}
void resnet_use_strict(ResNetDevice* r) {
    r->use_tc = false; r->split = false; r->fell_back = 1;
    r->state_elems = r->C * r->hh * r->hw;           // dense NCHW states: smaller than the board layout, the pool fits
}

// Stand-alone conv3x3 (+bias, +residual, +ReLU) on host NCHW data through either implementation.
// Debug / parity entry point behind mz_debug_conv3x3.
int resnet_debug_conv(int n, int C, int H, int W, const float* x, const float* w_oihw, const float* bias,
                      const float* residual, int relu, int use_tc, float* out, int sm_count, std::string* err) {
    if (use_tc && !conv_tc_supported(C, H, W)) { *err = "shape not supported by the tensor-core conv"; return MZ_EUNSUPPORTED; }
    if (C % 4) { *err = "C must be a multiple of 4"; return MZ_EINVAL; }
    MzNetDesc nd{};
    nd.kind = MZ_NET_RESNET; nd.channels = C; nd.obs_c = C; nd.obs_h = H; nd.obs_w = W; nd.action_space = 1;
    ResNetDevice r{};
    r.net = nd; r.max_batch = n; r.sm_count = sm_count; r.C = C; r.hh = H; r.hw = W;
    // fake a one-tensor state_dict for pack_conv
    MzTensor t{"conv.weight", w_oihw, (int64_t)C * C * 9};
    Loader L{&t, 1, err};
    std::vector<float> blob;
    std::vector<ConvLayer> layers;
    r.use_tc = use_tc != 0; r.split = use_tc == 2; r.tc_capable = true;
    if (!pack_conv(L, "conv", "", C, C, 1, blob, layers, use_tc == 2 ? kLayoutSplit : (use_tc ? kLayoutF16 : 0), H, W)) return MZ_EINVAL;
    long bias_off = -1;
    if (bias) { bias_off = (long)blob.size(); blob.insert(blob.end(), bias, bias + C); while (blob.size() % 4) blob.push_back(0.f); }
    layers[0].b_off = bias_off;
    const bool split = use_tc == 2;
    const size_t dense = (size_t)n * C * H * W, packed = (size_t)n * conv_tc_board_elems(split);
    float *d_blob = nullptr, *d_x = nullptr, *d_res = nullptr, *d_out = nullptr, *d_px = nullptr, *d_pres = nullptr, *d_pout = nullptr;
    auto cleanup = [&]() { for (float* p : {d_blob, d_x, d_res, d_out, d_px, d_pres, d_pout}) if (p) cudaFree(p); };
    bool ok = cudaMalloc(&d_blob, blob.size() * 4) == cudaSuccess && cudaMalloc(&d_x, dense * 4) == cudaSuccess &&
              cudaMalloc(&d_out, dense * 4) == cudaSuccess && (!residual || cudaMalloc(&d_res, dense * 4) == cudaSuccess);
    if (ok && use_tc)
        ok = cudaMalloc(&d_px, packed * 4) == cudaSuccess && cudaMalloc(&d_pout, packed * 4) == cudaSuccess &&
             (!residual || cudaMalloc(&d_pres, packed * 4) == cudaSuccess);
    if (!ok) { cleanup(); *err = "allocation failed"; return MZ_ENOMEM; }
    cudaMemcpy(d_blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_x, x, dense * 4, cudaMemcpyHostToDevice);
    if (residual) cudaMemcpy(d_res, residual, dense * 4, cudaMemcpyHostToDevice);
    r.d_conv = d_blob;
    int64_t launches = 0;
    Runner R{&r, nullptr, &launches, err, n};
    bool good;
    if (use_tc) {
        const unsigned blocks = (unsigned)((dense + 255) / 256);
        cudaMemset(d_px, 0, packed * 4);
        nchw_to_p64c4_kernel<<<blocks, 256>>>(d_x, d_px, n, C, H, W, split ? 1 : 0);
        if (residual) { cudaMemset(d_pres, 0, packed * 4); nchw_to_p64c4_kernel<<<blocks, 256>>>(d_res, d_pres, n, C, H, W, split ? 1 : 0); }
        good = R.conv_tc(layers[0], d_px, d_pout, residual ? d_pres : nullptr, relu != 0);
        if (good) p64c4_to_nchw_kernel<<<blocks, 256>>>(d_pout, d_out, n, C, H, W, split ? 1 : 0);
    } else {
        good = R.conv(layers[0], d_x, d_out, residual ? d_res : nullptr, relu != 0, H, W);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (good && e != cudaSuccess) { good = false; *err = std::string("debug conv: ") + cudaGetErrorString(e); }
    // optional warm-L2 timing of the bare kernel: MZ_DEBUG_CONV_REPS=k prints the mean of k back-to-back launches
    const char* reps_env = getenv("MZ_DEBUG_CONV_REPS");
    if (good && reps_env && atoi(reps_env) > 0) {
        const int reps = atoi(reps_env);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) {
            if (use_tc) R.conv_tc(layers[0], d_px, d_pout, residual ? d_pres : nullptr, relu != 0);
            else R.conv(layers[0], d_x, d_out, residual ? d_res : nullptr, relu != 0, H, W);
        }
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * n * H * W * (double)C * C * 9;
        fprintf(stderr, "[mz_debug_conv3x3] %s n=%d C=%d %dx%d residual=%d: %.2f us per launch, %.1f TFLOP/s useful\n",
                use_tc ? "tcgen05" : "cuda-core", n, C, H, W, residual ? 1 : 0, 1000.0 * ms / reps, flops / (ms / reps * 1e-3) / 1e12);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    if (good) cudaMemcpy(out, d_out, dense * 4, cudaMemcpyDeviceToHost);
    r.d_conv = nullptr;
    cleanup();
    return good ? MZ_OK : MZ_ECUDA;
}

// stored hidden states (pool layout) -> dense NCHW, device to device
int resnet_states_to_nchw(ResNetDevice* r, const float* states, int count, float* out, cudaStream_t stream) {
    const size_t total = (size_t)count * r->C * r->hh * r->hw;
    if (r->use_tc) p64c4_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(states, out, count, r->C, r->hh, r->hw, r->split ? 1 : 0);
    else cudaMemcpyAsync(out, states, total * 4, cudaMemcpyDeviceToDevice, stream);
    return cudaGetLastError() == cudaSuccess ? MZ_OK : MZ_ECUDA;
}

// dense NCHW states -> the pool layout, device to device (mz_import_tree)
int resnet_states_from_nchw(ResNetDevice* r, const float* dense, int count, float* states, cudaStream_t stream) {
    const size_t total = (size_t)count * r->C * r->hh * r->hw;
    if (r->use_tc) {
        cudaMemsetAsync(states, 0, (size_t)count * r->state_elems * 4, stream);          // padding positions read as zero
        nchw_to_p64c4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dense, states, count, r->C, r->hh, r->hw, r->split ? 1 : 0);
    } else {
        cudaMemcpyAsync(states, dense, total * 4, cudaMemcpyDeviceToDevice, stream);
    }
    return cudaGetLastError() == cudaSuccess ? MZ_OK : MZ_ECUDA;
}

int resnet_inference(ResNetDevice* r, const InferCall& c, cudaStream_t stream, int64_t* launches, std::string* err) {
    if (!r->loaded) { *err = "weights not loaded"; return MZ_ESTATE; }
    if (c.n > r->max_batch) { *err = "batch larger than max_games"; return MZ_EINVAL; }
    if (r->use_tc) return resnet_inference_tc(r, c, stream, launches, err);
    const MzNetDesc& nd = r->net;
    const int n = c.n, C = r->C, hh = r->hh, hw = r->hw, F = 2 * nd.support_size + 1;
    Runner R{r, stream, launches, err, n};
    float *cur = r->ws[0], *tmp = r->ws[1], *spare = r->ws[2];
    float* hidden_out = c.hidden ? c.hidden : r->scratch_hidden;

    if (!c.recurrent) {
        int H = nd.obs_h, W = nd.obs_w;
        if (nd.downsample) {
            const auto& d = r->rep_down;
            if (!R.conv(d[0], c.in, cur, nullptr, false, H, W)) return MZ_ECUDA;
            H = conv_out(H, 2); W = conv_out(W, 2);
            if (!R.blocks(d, 1, 2, &cur, &tmp, &spare, H, W)) return MZ_ECUDA;
            if (!R.conv(d[5], cur, tmp, nullptr, false, H, W)) return MZ_ECUDA;
            { float* t = cur; cur = tmp; tmp = t; }
            H = conv_out(H, 2); W = conv_out(W, 2);
            if (!R.blocks(d, 6, 3, &cur, &tmp, &spare, H, W)) return MZ_ECUDA;
            for (int pool = 0; pool < 2; ++pool) {
                const int Ho = conv_out(H, 2), Wo = conv_out(W, 2);
                const size_t total = (size_t)n * C * Ho * Wo;
                avgpool3x3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(cur, tmp, n * C, H, W, Ho, Wo);
                *launches += 1;
                { float* t = cur; cur = tmp; tmp = t; }
                H = Ho; W = Wo;
                if (pool == 0 && !R.blocks(d, 12, 3, &cur, &tmp, &spare, H, W)) return MZ_ECUDA;
            }
            const int fused = R.small_tower(r->rep_trunk, 0, false, nd.blocks, cur, tmp, C, H, W);
            if (fused < 0) return MZ_ECUDA;
            if (fused) { float* t = cur; cur = tmp; tmp = t; }
            else if (!R.blocks(r->rep_trunk, 0, nd.blocks, &cur, &tmp, &spare, H, W)) return MZ_ECUDA;
        } else {
            const int fused = R.small_tower(r->rep_trunk, 0, true, nd.blocks, c.in, cur, nd.obs_c, H, W);
            if (fused < 0) return MZ_ECUDA;
            if (!fused) {
                if (!R.conv(r->rep_trunk[0], c.in, cur, nullptr, true, H, W)) return MZ_ECUDA;
                if (!R.blocks(r->rep_trunk, 1, nd.blocks, &cur, &tmp, &spare, H, W)) return MZ_ECUDA;
            }
        }
        // rescale -> hidden (no heads on the raw state at the root)
        if (!R.heads(cur, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, hidden_out, c.pool_hidden, c.pool_stride, c.out_slot))
            return MZ_ECUDA;
        if (c.reward_logits) {
            fill_root_reward_logits_kernel<<<(n * F + 255) / 256, 256, 0, stream>>>(c.reward_logits, n, F, nd.support_size);
            *launches += 1;
        }
        if (c.reward) {
            fill_root_reward_kernel<<<(n + 255) / 256, 256, 0, stream>>>(c.reward, n);
            *launches += 1;
        }
    } else {
        const float* in = c.gather_parent ? c.pool_hidden : c.in;
        const int fused = R.small_tower(r->dyn, 0, true, nd.blocks, in, cur, C, hh, hw, c.gather_parent, c.pool_stride, c.action);
        if (fused < 0) return MZ_ECUDA;
        if (!fused) {
            if (!R.conv(r->dyn[0], in, cur, nullptr, true, hh, hw, c.gather_parent, c.pool_stride, c.action)) return MZ_ECUDA;
            if (!R.blocks(r->dyn, 1, nd.blocks, &cur, &tmp, &spare, hh, hw)) return MZ_ECUDA;
        }
        // reward head on the raw state + rescale -> hidden
        if (!R.heads(cur, 1, &r->reward_head, nullptr, c.reward_logits, nullptr, c.reward, nullptr, hidden_out, c.pool_hidden,
                     c.pool_stride, c.out_slot))
            return MZ_ECUDA;
    }
    // prediction on the rescaled state
    {
        float* x = hidden_out;
        // the tower must not overwrite the hidden state: first conv reads it, writes workspace
        float *pc = cur, *pt = tmp, *ps = spare;
        const int fused = nd.blocks > 0 ? R.small_tower(r->pred, 0, false, nd.blocks, x, pt, C, hh, hw) : 0;
        if (fused < 0) return MZ_ECUDA;
        if (fused) {
            x = pt;
        } else if (nd.blocks > 0) {
            if (!R.conv(r->pred[0], x, pt, nullptr, true, hh, hw)) return MZ_ECUDA;
            if (!R.conv(r->pred[1], pt, ps, x, true, hh, hw)) return MZ_ECUDA;
            { float* t = pc; pc = ps; ps = t; }
            if (!R.blocks(r->pred, 2, nd.blocks - 1, &pc, &pt, &ps, hh, hw)) return MZ_ECUDA;
            x = pc;
        }
        if (!R.heads(x, 2, &r->value_head, &r->policy_head, c.value_logits, c.policy_logits, c.value, nullptr, nullptr, nullptr, 0, 0))
            return MZ_ECUDA;
    }
    return MZ_OK;
}

}  // namespace mz
