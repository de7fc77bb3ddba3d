// Shared device helpers: lane groups, reductions, Philox4x32-10, fp32 scalarisation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MZ_DEVINL __device__ __forceinline__

namespace mz {

// ------------------------------------------------------------------------------------------
// Lane groups: a warp is split into 32/G groups of G consecutive lanes; one game per group.
// ------------------------------------------------------------------------------------------
template <int G>
struct LaneGroup {
    static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group width");
    MZ_DEVINL static unsigned lane() { return threadIdx.x & (G - 1); }
    MZ_DEVINL static unsigned base() { return (threadIdx.x & 31u) & ~(unsigned)(G - 1); }
    MZ_DEVINL static unsigned mask() {
        if constexpr (G == 32) return 0xffffffffu;
        else return ((1u << G) - 1u) << base();
    }
    MZ_DEVINL static void sync() { __syncwarp(mask()); }
    // ballot restricted to the group, bit i = lane i of the group
    MZ_DEVINL static unsigned ballot(bool p) {
        unsigned b = __ballot_sync(mask(), p);
        if constexpr (G == 32) return b;
        else return (b >> base()) & ((1u << G) - 1u);
    }
    template <typename T>
    MZ_DEVINL static T bcast(T v, int src) { return __shfl_sync(mask(), v, src, G); }
};

MZ_DEVINL double shfl_xor_f64(unsigned mask, double v, int off, int width) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(mask, lo, off, width);
    hi = __shfl_xor_sync(mask, hi, off, width);
    return __hiloint2double(hi, lo);
}
MZ_DEVINL double shfl_f64(unsigned mask, double v, int src, int width) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(mask, lo, src, width);
    hi = __shfl_sync(mask, hi, src, width);
    return __hiloint2double(hi, lo);
}

// smallest power of two >= n (n >= 1)
// d / sc (IEEE round-to-nearest) for sc > 0.  A zero numerator - every channel minimum, every ReLU zero - sends
// div.rn.f32 down its out-of-line slow path (FCHK rejects zero / denormal operands) and the whole warp waits for it;
// 0 / sc is +0 anyway, so divide a harmless 1.0 instead and select.
MZ_DEVINL float div_pos_or_zero(float d, float sc) {
    const float q = __fdiv_rn(d == 0.0f ? 1.0f : d, sc);
    return d == 0.0f ? 0.0f : q;
}

// hint: bring the line holding *p into L1 (no register, no dependency)
MZ_DEVINL void prefetch_l1(const void* p) { asm volatile("prefetch.L1 [%0];" ::"l"(p)); }

MZ_DEVINL int pow2_ceil(int n) { return n <= 1 ? 1 : 1 << (32 - __clz(n - 1)); }

// Exact max / min (no rounding involved), NaN-free inputs assumed.
template <int G>
MZ_DEVINL double group_max_f64(double v, int width) {
    const unsigned m = LaneGroup<G>::mask();
    for (int off = width >> 1; off > 0; off >>= 1) v = fmax(v, shfl_xor_f64(m, v, off, G));
    return v;
}
template <int G>
MZ_DEVINL double group_min_f64(double v, int width) {
    const unsigned m = LaneGroup<G>::mask();
    for (int off = width >> 1; off > 0; off >>= 1) v = fmin(v, shfl_xor_f64(m, v, off, G));
    return v;
}
template <int G>
MZ_DEVINL float group_max_f32(float v) {
    const unsigned m = LaneGroup<G>::mask();
#pragma unroll
    for (int off = G >> 1; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(m, v, off, G));
    return v;
}
// The same reductions over the first W lanes of the group only (W a power of two <= G): valid in lanes < W.  With the lanes
// >= W holding the neutral element the full-width reduction returns the same bits (max is exact, x + 0 is exact): these
// just skip the steps that cannot change the result.
constexpr int pow2_ceil_c(int n) { return n <= 1 ? 1 : 2 * pow2_ceil_c((n + 1) / 2); }
template <int G, int W>
MZ_DEVINL float group_max_f32_w(float v) {
    const unsigned m = LaneGroup<G>::mask();
#pragma unroll
    for (int off = W >> 1; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(m, v, off, G));
    return v;
}
template <int G, int W>
MZ_DEVINL float group_sum_f32_w(float v) {
    const unsigned m = LaneGroup<G>::mask();
#pragma unroll
    for (int off = W >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(m, v, off, G);
    return v;
}
template <int G>
MZ_DEVINL float group_sum_f32(float v) {
    const unsigned m = LaneGroup<G>::mask();
#pragma unroll
    for (int off = G >> 1; off > 0; off >>= 1) v += __shfl_xor_sync(m, v, off, G);
    return v;
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11); mirrored in oracle/philox.py.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;
constexpr uint32_t kTagTie = 0x7169E001u, kTagNoise = 0x7169E002u, kTagAction = 0x7169E003u;

struct Philox4 { uint32_t x, y, z, w; };

MZ_DEVINL Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t p0h = __umulhi(kPhiloxM0, c0), p0l = kPhiloxM0 * c0;
        const uint32_t p1h = __umulhi(kPhiloxM1, c2), p1l = kPhiloxM1 * c2;
        c0 = p1h ^ c1 ^ k0; c1 = p1l;
        c2 = p0h ^ c3 ^ k1; c3 = p0l;
        k0 += kPhiloxW0; k1 += kPhiloxW1;
    }
    return Philox4{c0, c1, c2, c3};
}

// index in [0, n) for an exact UCB tie at (game, move, sim, depth)
MZ_DEVINL int philox_tie_index(uint64_t seed, int64_t game, int move, int sim, int depth, int n) {
    const Philox4 r = philox4x32_10((uint32_t)game, (uint32_t)move, (uint32_t)sim, (uint32_t)depth,
                                    (uint32_t)seed, (uint32_t)(seed >> 32) ^ kTagTie);
    return (int)__umulhi(r.x, (uint32_t)n);
}

// Gamma(alpha, 1) sample for lane-private use (Marsaglia & Tsang 2000, with the alpha < 1 boost
// gamma(alpha) = gamma(alpha+1) * U^(1/alpha)); uniforms from Philox keyed (game, move, lane, draw).
// Used to draw the root Dirichlet noise on the device (self_play.py:473) when the host passes none.
static __device__ __noinline__ double philox_gamma(uint64_t seed, int64_t game, int move, int lane, double alpha) {
    const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    const double k2 = 1.0 / 4294967296.0;
    double out = d;
    for (int it = 0; it < 64; ++it) {
        const Philox4 r = philox4x32_10((uint32_t)game, (uint32_t)move, (uint32_t)lane, (uint32_t)it,
                                        (uint32_t)seed, (uint32_t)(seed >> 32) ^ kTagNoise);
        const double u1 = ((double)r.x + 0.5) * k2, u2 = ((double)r.y + 0.5) * k2, u3 = ((double)r.z + 0.5) * k2;
        const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);          // Box-Muller
        const double t = 1.0 + c * x;
        if (t <= 0.0) continue;
        const double v = t * t * t;
        if (log(u3) < 0.5 * x * x + d - d * v + d * log(v)) { out = d * v; break; }
    }
    if (alpha < 1.0) {
        const Philox4 r = philox4x32_10((uint32_t)game, (uint32_t)move, (uint32_t)lane, 0xFFFFu,
                                        (uint32_t)seed, (uint32_t)(seed >> 32) ^ kTagNoise);
        out *= pow(((double)r.x + 0.5) * k2, 1.0 / alpha);
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// fp32 helpers written with explicit rounding so -fmad cannot change them.
// ------------------------------------------------------------------------------------------
// models.py:661-665 applied to x = sum_k k * softmax(logits)_k
MZ_DEVINL float inverse_value_transform(float x) {
    const float eps = 0.001f;
    const float ax = fabsf(x);
    float t = __fadd_rn(__fadd_rn(ax, 1.0f), eps);          // |x| + 1 + 0.001
    t = __fadd_rn(1.0f, __fmul_rn(4.0f * eps, t));          // 1 + 4*0.001*(...)
    t = __fsub_rn(__fsqrt_rn(t), 1.0f);                     // sqrt(...) - 1
    t = __fdiv_rn(t, 2.0f * eps);                           // / (2*0.001)
    t = __fsub_rn(__fmul_rn(t, t), 1.0f);                   // **2 - 1
    const float sgn = (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);   // torch.sign
    return __fmul_rn(sgn, t);
}

MZ_DEVINL float elu1(float x) { return x > 0.0f ? x : (expf(x) - 1.0f); }

}  // namespace mz
