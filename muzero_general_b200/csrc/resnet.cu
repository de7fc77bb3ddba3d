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
#include "heads.cuh"
#include "ktimer.h"
#include "small_tower.h"
#include "small_search.h"
#include "launch.h"
#include "conv_tc.h"
#include "pipeline.h"

namespace mz {

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

template <int GROUP>
__global__ void __launch_bounds__(kHeadThreads) heads_kernel(const __grid_constant__ HeadsArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int group = threadIdx.x / GROUP, t = threadIdx.x % GROUP, ngroups = blockDim.x / GROUP;
    float* s_w = sm;                                                 // head blob slice [w_lo, w_lo + w_floats)
    float* scratch = s_w + a.w_floats + (size_t)group * a.warp_floats;
    // board layout only (C = 64): padded row of dense position p, looked up instead of two runtime divisions per chunk
    __shared__ unsigned char s_pos[64];
    pdl_launch_dependents();
    {
        const float4* src = reinterpret_cast<const float4*>(a.blob + a.w_lo);
        float4* dst = reinterpret_cast<float4*>(s_w);
        for (int i = threadIdx.x; i < a.w_floats / 4; i += blockDim.x) dst[i] = src[i];
        if (a.p64c4 && threadIdx.x < a.HW && threadIdx.x < 64)
            s_pos[threadIdx.x] = (unsigned char)((threadIdx.x / a.W + 1) * 8 + (threadIdx.x % a.W));
    }
    __syncthreads();
    pdl_wait();                                                      // the weights are constants; x comes from the previous kernel
    const float* blob = s_w - a.w_lo;                                // blob[off] addresses the staged copy
    for (int g = a.g0 + blockIdx.x * ngroups + group; g < a.g0 + a.n; g += gridDim.x * ngroups)
        heads_one_sample<GROUP>(a, blob, scratch, s_pos, g, group, t, a.out_slot);
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
    int g0 = 0;                // games [g0, g0 + n) of the batch (partitioned replay); buffers are addressed by the global index
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
        a.g0 = g0; a.sat_count = r->d_sat;
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
        if (g0 != 0) { *err = "conv3x3: partitioned calls are not supported on the per-layer route"; return false; }
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
    // arguments of a fused CUDA-core tower; false when the layers are not of the supported kind
    bool small_tower_args(SmallTowerArgs& a, const std::vector<ConvLayer>& layers, size_t first, bool stem, size_t count, const float* in,
                          float* out, int in_channels, int H, int W, const int32_t* gather_parent, int pool_stride, const int32_t* action) {
        const size_t nl = (stem ? 1 : 0) + 2 * count;
        if (nl == 0 || nl > (size_t)kSmallTowerMaxLayers) return false;
        a = SmallTowerArgs{};
        a.in = in; a.out = out; a.blob = r->d_conv; a.gather_parent = gather_parent; a.action = action; a.pool_stride = pool_stride;
        a.n = n; a.g0 = g0; a.C = r->C; a.H = H; a.W = W; a.A = r->net.action_space; a.in_channels = in_channels; a.n_layers = (int)nl;
        for (size_t i = 0; i < nl; ++i) {
            const ConvLayer& l = layers[first + i];
            if (l.stride != 1 || l.cout != r->C) return false;
            SmallTowerLayer& t = a.layer[i];
            t.w_off = (int)l.w_off; t.b_off = (int)l.b_off; t.cin = l.cin; t.relu = 1;
            t.residual = (i >= (stem ? 1u : 0u) && ((i - (stem ? 1 : 0)) & 1)) ? 1 : 0;     // second conv of a block
        }
        return true;
    }

    int small_tower(const std::vector<ConvLayer>& layers, size_t first, bool stem, size_t count, const float* in, float* out,
                    int in_channels, int H, int W, const int32_t* gather_parent = nullptr, int pool_stride = 0,
                    const int32_t* action = nullptr, bool dry_run = false) {
        if (!r->fuse_small) return 0;
        SmallTowerArgs a{};
        if (!small_tower_args(a, layers, first, stem, count, in, out, in_channels, H, W, gather_parent, pool_stride, action)) return 0;
        if (!small_tower_supported(a)) return 0;
        if (dry_run) return 1;
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
        if (g0 != 0) { *err = "heads (generic route): partitioned calls are not supported"; return false; }
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

    // arguments of one heads launch (everything but the launch geometry)
    HeadsArgs heads_args(const float* x, int n_heads, const HeadDesc* h0, const HeadDesc* h1, float* l0, float* l1, float* s0, float* s1,
                         float* rescaled, float* pool_hidden, int pool_stride, int out_slot, bool p64c4 = false, float* state_p64c4 = nullptr) {
        HeadsArgs a{};
        a.p64c4 = p64c4 ? (r->split ? kLayoutSplit : kLayoutF16) : 0; a.W = r->hw; a.state_p64c4 = state_p64c4;
        a.x = x; a.blob = r->d_head; a.n = n; a.g0 = g0; a.C = r->C; a.HW = r->hh * r->hw; a.S = r->net.support_size;
        a.hw_inv = a.HW >= 2 ? (unsigned)((0x100000000ull + (unsigned)a.HW - 1u) / (unsigned)a.HW) : 0u;
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
        return a;
    }

    bool heads(const float* x, int n_heads, const HeadDesc* h0, const HeadDesc* h1, float* l0, float* l1, float* s0, float* s1,
               float* rescaled, float* pool_hidden, int pool_stride, int out_slot, bool p64c4 = false, float* state_p64c4 = nullptr) {
        HeadsArgs a = heads_args(x, n_heads, h0, h1, l0, l1, s0, s1, rescaled, pool_hidden, pool_stride, out_slot, p64c4, state_p64c4);
        const HeadDesc* hs[2] = {h0, h1};
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
    Runner R{r, stream, launches, err, n, c.g0};
    if (c.g0 != 0 && (!r->split || !c.recurrent || !c.gather_parent)) { *err = "resnet: partitioned calls need the x3 towers in pool mode"; return MZ_EINVAL; }
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
bool resnet_uses_tensor_cores(const ResNetDevice* r) { return r->use_tc; }

// Partitioned replay (abi.cu) needs every kernel of a recurrent inference to honour a first-game offset: the x3 towers,
// the fused small towers and heads_kernel do; the per-layer convs, the fp16-mode towers and the generic heads route do not.
static const int32_t kDryRunAction = 0;        // stands for the action array in dry runs (only its presence matters)
bool resnet_can_partition(const ResNetDevice* r0) {
    ResNetDevice* r = const_cast<ResNetDevice*>(r0);
    if (!r->loaded) return false;
    if (r->use_tc) return r->split;
    std::string err; int64_t launches = 0;
    Runner R{r, nullptr, &launches, &err, r->max_batch, 0};
    const int nb = r->net.blocks;
    if (nb < 1) return false;
    if (R.small_tower(r->dyn, 0, true, nb, nullptr, nullptr, r->C, r->hh, r->hw, nullptr, 0, &kDryRunAction, true) != 1) return false;
    if (R.small_tower(r->pred, 0, false, nb, nullptr, nullptr, r->C, r->hh, r->hw, nullptr, 0, nullptr, true) != 1) return false;
    // heads_kernel route (not heads_big): the head weights plus one group's tile fit in shared memory
    const int HW = r->hh * r->hw;
    int hi = 0, lo = 1 << 30, maxw = 32;
    for (const HeadDesc* d : {&r->reward_head, &r->value_head, &r->policy_head}) {
        lo = std::min(lo, d->w1_off);
        hi = std::max(hi, d->mlp.b_off[d->mlp.n - 1] + d->mlp.out[d->mlp.n - 1]);
        maxw = std::max(maxw, d->rc * HW + 4);
        for (int l = 0; l < d->mlp.n; ++l) maxw = std::max(maxw, d->mlp.out[l] + 4);
    }
    const size_t warp_floats = (size_t)HW * (r->C + 4) + 6 * r->C + 4 * ((maxw + 3) & ~3);
    return ((size_t)(hi - lo) + warp_floats) * 4 <= 227 * 1024;
}
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

// Fused search for small residual networks (small_search.cu): all the simulations of the games [c.g0, c.g0 + c.n) in ONE
// launch.  `c` is the recurrent call of the first simulation (pool mode), `tree` the tree step that follows it; the
// arguments of the towers and the heads are the ones resnet_inference would launch with, simulation after simulation.
static bool small_search_build(ResNetDevice* r, const InferCall& c, const TreeStepArgs& tree, int n_sims, SmallSearchArgs* out,
                               int* P, int* CO, int* G, int* threads, size_t* smem) {
    if (r->use_tc || !r->loaded || !r->fuse_small || !c.recurrent || !c.gather_parent || c.hidden || r->net.blocks < 1) return false;
    if (c.value_logits || c.reward_logits) return false;
    std::string err; int64_t launches = 0;
    Runner R{r, nullptr, &launches, &err, c.n, c.g0};
    const int C = r->C, hh = r->hh, hw = r->hw, nb = r->net.blocks;
    SmallSearchArgs a{};
    float* raw = r->ws[0];                 // dynamics tower output
    float* pred_out = r->ws[1];            // prediction tower output
    float* hidden = r->scratch_hidden;     // rescaled state, dense (input of the prediction tower)
    if (!R.small_tower_args(a.dyn, r->dyn, 0, true, nb, c.pool_hidden, raw, C, hh, hw, c.gather_parent, c.pool_stride, c.action)) return false;
    if (!R.small_tower_args(a.pred, r->pred, 0, false, nb, hidden, pred_out, C, hh, hw, nullptr, 0, nullptr)) return false;
    if (!small_tower_layout(a.dyn) || !small_tower_layout(a.pred)) return false;
    const int cap = std::max(a.dyn.cap_channels, a.pred.cap_channels);
    a.dyn.cap_channels = a.pred.cap_channels = cap;
    a.heads_dyn = R.heads_args(raw, 1, &r->reward_head, nullptr, nullptr, nullptr, c.reward, nullptr, nullptr, c.pool_hidden, c.pool_stride, c.out_slot);
    a.heads_pred = R.heads_args(pred_out, 2, &r->value_head, &r->policy_head, nullptr, c.policy_logits, c.value, nullptr, nullptr, nullptr, 0, 0);
    if (!(a.heads_dyn.C * a.heads_dyn.HW <= 1024)) return false;                   // one warp per sample (heads_kernel<32>)
    const int lo = std::min(a.heads_dyn.w_lo, a.heads_pred.w_lo);
    const int hi = std::max(a.heads_dyn.w_lo + a.heads_dyn.w_floats, a.heads_pred.w_lo + a.heads_pred.w_floats);
    a.heads_lo = lo & ~3; a.heads_floats = ((hi - a.heads_lo) + 3) & ~3;
    a.scratch_floats = std::max(a.heads_dyn.warp_floats, a.heads_pred.warp_floats);
    a.tree = tree;
    a.n = c.n; a.g0 = c.g0; a.n_sims = n_sims; a.first_slot = c.out_slot;
    int tile = 0, row_stride = 0, board_stride = 0;
    const int tower_floats = ((a.dyn.w_floats + 3) & ~3) + ((a.pred.w_floats + 3) & ~3);
    if (!small_search_shape(hh, hw, C, r->net.action_space, c.n, r->sm_count, tower_floats, a.heads_floats, a.scratch_floats, cap,
                            P, CO, G, &tile, threads, smem, &row_stride, &board_stride))
        return false;
    a.tile = tile;
    a.dyn.boards_per_cta = a.pred.boards_per_cta = tile;
    a.dyn.row_stride = a.pred.row_stride = row_stride;
    a.dyn.board_stride = a.pred.board_stride = board_stride;
    a.off_wd = 0;
    a.off_wp = (a.dyn.w_floats + 3) & ~3;
    a.off_wh = tower_floats;
    a.off_scratch = a.off_wh + a.heads_floats;
    a.off_map = a.off_scratch + (*threads / 32) * a.scratch_floats;
    a.off_act = a.off_map + 2 * C * hh * hw;
    *out = a;
    return true;
}

bool resnet_small_search_supported(ResNetDevice* r, const InferCall& c, const TreeStepArgs& tree, int n_sims) {
    // A/B switch: MZ_SMALL_SEARCH=0 keeps the step-wise pipeline, =1 uses the fused kernel wherever the shape allows
    constexpr bool kDefaultOn = true;
    const char* sw = getenv("MZ_SMALL_SEARCH");
    if (sw ? sw[0] != '1' : !kDefaultOn) return false;
    SmallSearchArgs a; int P, CO, G, threads; size_t smem;
    return small_search_build(r, c, tree, n_sims, &a, &P, &CO, &G, &threads, &smem);
}

int resnet_small_search(ResNetDevice* r, const InferCall& c, const TreeStepArgs& tree, int n_sims, cudaStream_t stream, int64_t* launches,
                        std::string* err) {
    SmallSearchArgs a; int P, CO, G, threads; size_t smem;
    if (!small_search_build(r, c, tree, n_sims, &a, &P, &CO, &G, &threads, &smem)) { *err = "small_search: shape not supported"; return MZ_EINVAL; }
    kt_begin(KT_SEARCH, stream);
    cudaError_t e = launch_small_search(a, P, CO, G, threads, smem, stream);
    kt_end(stream);
    if (e != cudaSuccess) { *err = std::string("small_search launch: ") + cudaGetErrorString(e); return MZ_ECUDA; }
    *launches += 1;
    return MZ_OK;
}

int resnet_inference(ResNetDevice* r, const InferCall& c, cudaStream_t stream, int64_t* launches, std::string* err) {
    if (!r->loaded) { *err = "weights not loaded"; return MZ_ESTATE; }
    if (c.g0 < 0 || c.g0 + c.n > r->max_batch) { *err = "batch larger than max_games"; return MZ_EINVAL; }
    if (r->use_tc) return resnet_inference_tc(r, c, stream, launches, err);
    const MzNetDesc& nd = r->net;
    const int n = c.n, C = r->C, hh = r->hh, hw = r->hw, F = 2 * nd.support_size + 1;
    Runner R{r, stream, launches, err, n, c.g0};
    if (c.g0 != 0 && !c.recurrent) { *err = "resnet: partitioned calls are recurrent only"; return MZ_EINVAL; }
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
