// Issue-rate probe for the tcgen05 shapes conv_x3.cu / conv_tc.cu use (profiles/r02_mma_rate.md).
// One CTA per SM, one elected thread issues `layers` x 36 K-steps of a pattern back to back (no epilogue, nothing waits
// besides the accumulator dependencies), commits once and waits; clock64 around it.  Operands are the SAME shared-memory
// images the tower kernel uses: two 36 KB activation planes (K-major, SWIZZLE_128B, 128-byte rows) and a 144 KB weight
// image [9 taps][128 rows][64 cin].  A pattern is a list of up to four MMAs per K-step.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I muzero_general_b200/csrc -o scripts/mma_rate scripts/mma_rate.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace mz::tc;

constexpr int kRowBytes = 128, kHalo = 16, kRowsX = 288, kPlaneBytes = kRowsX * kRowBytes, kTapBytes = 128 * kRowBytes;
constexpr int kWBytes = 9 * kTapBytes;
constexpr int kSmem = kWBytes + 2 * kPlaneBytes + 1024;

__device__ __forceinline__ void umma(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc_word, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc_word), "r"(acc), "r"(kDescHi) : "memory");
}

__host__ __device__ constexpr uint32_t idesc(uint32_t n, uint32_t m = 128) { return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24); }

enum Pat {
    kX3, kN128, kN64, kN256, kN128x2, kX3BothTiles, kN64x3, kN128TwoAcc, kX3Reversed, kN128Overwrite, kX3NoShift, kN32, kN16, kPatterns
};

// The 36 K-steps of one layer, fully unrolled with compile-time offsets (the shape of the tower kernel's issue loop).
template <int PAT>
__device__ __forceinline__ void issue_layer(uint32_t tmem, uint32_t ah, uint32_t al, uint32_t w16) {
    uint32_t acc = 0;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int shift = PAT == kX3NoShift ? 0 : (tap / 3 - 1) * 8 + (tap % 3 - 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const uint32_t off = (uint32_t)(shift * 8 + ks * 2);
            const uint32_t b = w16 + (uint32_t)(tap * (kTapBytes / 16) + ks * 2);
            const uint32_t b2 = w16 + (uint32_t)((tap & 6) * (kTapBytes / 16) + ks * 2);      // room for 256 rows
            const uint32_t t1 = (uint32_t)(128 * kRowBytes / 16);                              // second tile: +128 rows
            if constexpr (PAT == kX3 || PAT == kX3NoShift) { umma(tmem, ah + off, b, idesc(128), acc); umma(tmem + 128, al + off, b, idesc(64), acc); }
            if constexpr (PAT == kN128) umma(tmem, ah + off, b, idesc(128), acc);
            if constexpr (PAT == kN64) umma(tmem, ah + off, b, idesc(64), acc);
            if constexpr (PAT == kN32) umma(tmem, ah + off, b, idesc(32), acc);
            if constexpr (PAT == kN16) umma(tmem, ah + off, b, idesc(16), acc);
            if constexpr (PAT == kN256) umma(tmem, ah + off, b2, idesc(256), acc);
            if constexpr (PAT == kN128x2) { umma(tmem, ah + off, b, idesc(128), acc); umma(tmem + 128, al + off, b, idesc(128), acc); }
            if constexpr (PAT == kX3BothTiles) {
                umma(tmem, ah + off, b, idesc(128), acc); umma(tmem + 128, al + off, b, idesc(64), acc);
                umma(tmem + 256, ah + t1 + off, b, idesc(128), acc); umma(tmem + 384, al + t1 + off, b, idesc(64), acc);
            }
            if constexpr (PAT == kN64x3) {
                umma(tmem, ah + off, b, idesc(64), acc); umma(tmem + 64, ah + off, b + 64 * 8, idesc(64), acc);
                umma(tmem + 128, al + off, b, idesc(64), acc);
            }
            if constexpr (PAT == kN128TwoAcc) umma(tmem + (ks & 1) * 128, ah + off, b, idesc(128), (tap == 0 && ks < 2) ? 0u : 1u);
            if constexpr (PAT == kX3Reversed) { umma(tmem + 128, al + off, b, idesc(64), acc); umma(tmem, ah + off, b, idesc(128), acc); }
            if constexpr (PAT == kN128Overwrite) umma(tmem, ah + off, b, idesc(128), 0u);
            acc = 1;
        }
    }
}

template <int PAT>
__global__ void __launch_bounds__(128, 1) probe(int layers, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_w = s_base, s_hi = s_base + kWBytes, s_lo = s_hi + kPlaneBytes, bar = s_lo + kPlaneBytes;
    __shared__ uint32_t tmem_slot;
    // small finite fp16 values everywhere: the tensor core sees ordinary data
    for (int i = threadIdx.x; i < (kWBytes + 2 * kPlaneBytes) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = 0x2c002c00u + (uint32_t)(i * 2654435761u >> 28);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x < 32) {
        if (elect_one()) {
            const uint32_t w16 = (s_w >> 4) | kDescLoFlags;
            const uint32_t row0 = (uint32_t)(kHalo * kRowBytes);
            const uint32_t ah = ((s_hi + row0) >> 4) | kDescLoFlags, al = ((s_lo + row0) >> 4) | kDescLoFlags;
            const long long t0 = clock64();
            for (int l = 0; l < layers; ++l) issue_layer<PAT>(tmem, ah, al, w16);
            const long long t1 = clock64();
            umma_commit(bar);
            mbar_wait(bar, 0);
            cycles[2 * blockIdx.x] = clock64() - t0;
            cycles[2 * blockIdx.x + 1] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

struct Case { const char* name; void (*kern)(int, long long*); int mmas; double floor_cyc; };

int main(int argc, char** argv) {
    const int layers = argc > 1 ? atoi(argv[1]) : 40;
    const Case cases[] = {
        {"x3: N128 hi->D0 + N64 lo->D128 (the tower kernel's K-step)", probe<kX3>, 2, 96},
        {"x3, unshifted tap windows", probe<kX3NoShift>, 2, 96},
        {"x3 reversed: N64 first", probe<kX3Reversed>, 2, 96},
        {"N128", probe<kN128>, 1, 64},
        {"N128, overwrite instead of accumulate", probe<kN128Overwrite>, 1, 64},
        {"N128, two accumulators alternating", probe<kN128TwoAcc>, 1, 64},
        {"N64 (the fp16-mode kernel's K-step)", probe<kN64>, 1, 32},
        {"N32", probe<kN32>, 1, 16},
        {"N16", probe<kN16>, 1, 8},
        {"N256", probe<kN256>, 1, 128},
        {"2 x N128: hi->D0, lo->D128", probe<kN128x2>, 2, 128},
        {"3 x N64: hi w_h, hi w_l, lo w_h", probe<kN64x3>, 3, 96},
        {"x3 of both tiles per K-step (4 MMAs)", probe<kX3BothTiles>, 4, 192},
    };
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    long long* d_cycles;
    cudaMalloc(&d_cycles, sizeof(long long) * 2048);
    printf("| MMAs per K-step (M = 128, K = 16, fp16, SS) | cycles per K-step: 1 CTA | issue only | all %d SMs (median / max) | tensor floor |\n|---|---|---|---|---|\n",
           prop.multiProcessorCount);
    for (const Case& c : cases) {
        cudaFuncSetAttribute(c.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        double one = 0, issue = 0, med = 0, mx = 0;
        for (int ctas : {1, prop.multiProcessorCount}) {
            std::vector<long long> h(2 * ctas), tot(ctas);
            for (int rep = 0; rep < 2; ++rep) {
                c.kern<<<ctas, 128, kSmem>>>(layers, d_cycles);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h.data(), d_cycles, sizeof(long long) * 2 * ctas, cudaMemcpyDeviceToHost);
            for (int i = 0; i < ctas; ++i) tot[i] = h[2 * i];
            std::sort(tot.begin(), tot.end());
            const double steps = 36.0 * layers;
            if (ctas == 1) { one = tot[0] / steps; issue = h[1] / steps; }
            else { med = tot[ctas / 2] / steps; mx = tot[ctas - 1] / steps; }
        }
        printf("| %s | %.1f | %.1f | %.1f / %.1f | %.0f |\n", c.name, one, issue, med, mx, c.floor_cyc);
    }
    return 0;
}
