// conv3x3 (C -> C, stride 1, pad 1) as an implicit GEMM on the 5th-generation tensor cores.
//
// Used for the residual towers of board-sized states (H <= 6, W <= 7, C = 64: Connect4,
// models.py:213-229 inside representation / dynamics / prediction).  Two kernels share the MMA loop:
//
//   conv_tower_resident_kernel   up to 4 tiles per CTA (1184 boards on 148 SMs): the activations of a CTA's tiles stay
//                                in shared memory through all layers of a tower (see the comment above the kernel)
//   conv_tower_tc_kernel         larger batches / single convs: activations stream through L2, per CTA
//
//     weights  [tap 9][cout C][cin C] fp16, BN folded, 128B-swizzled   shared memory (bulk copy), two slot sets
//     A tile   two boards = 128 rows of the "P64S" layout    2-stage ring, one 8 KB cp.async.bulk per board
//     D        128 x 64 fp32 accumulator                     TMEM, double buffered
//     out      two output tiles staged in shared memory      one 8 KB bulk store per board (dedicated warp)
//
// Operands are fp16 (10-bit mantissa - the same as tf32 - with fp32 accumulation): one tcgen05.mma
// consumes K = 16 channels per 32-byte operand row, so a tile needs 36 MMAs instead of the 72 a tf32
// formulation needs, and every activation / weight byte moved through L2 and shared memory is halved.
// Measured motivation: profiles/r01_conv_tc_bottleneck.md (the M128 x N64 MMA is operand-fetch bound).
//
// P64S activation layout (HBM and shared): a board is 64 positions p = (y+1)*8 + x (row 0, rows H+1.. and columns
// W..7 are zero padding), every position one 128-byte row of 64 fp16 channels whose eight 16-byte chunks are stored
// XOR-ed with p % 8 - i.e. the boards sit in HBM already in the UMMA K-major SWIZZLE_128B shared-memory image, so a
// plain 1-D bulk copy lands them ready for the tensor core.  Filter tap (dy,dx) is the SAME shared-memory tile with
// its start address moved by (dy*8+dx) rows (the hardware swizzles on absolute address bits, so base_offset stays 0):
// the implicit GEMM needs no im2col copy.  Epilogue warps read the accumulator with tcgen05.ld, add the folded-BN
// bias, the optional residual and the optional action-plane term (models.py:557-572 folded into a per-position
// table), apply ReLU, zero the padding positions, convert to fp16 (round to nearest, saturating) and store P64S again.
//
// Warp roles (384 threads): 0 = bulk-copy producer, 1 = MMA issuer (one elected thread issues every tcgen05.mma of
// the CTA), 2 = TMEM allocator, 3 = output store (bulk copies shared -> global), 4..11 = epilogue (TMEM lane quarter
// = warp % 4; streaming kernel: accumulator column half = (warp - 4) / 4, resident kernel: tile parity = (warp - 4) / 4).
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include "pipeline.h"
#include "conv_tc.h"
#include "launch.h"
#include "tc_common.cuh"

namespace mz {

namespace {

using namespace tc;

constexpr int kC = 64;                 // channels in = out
constexpr int kPos = 64;               // positions per board (8 x 8 padded grid)
constexpr int kBoards = 2;             // boards per tile -> M = 128
constexpr int kHalo = 16;              // zero rows above / below the tile (|shift| <= 9; multiple of 8 keeps the swizzle phase)
constexpr int kRows = kBoards * kPos + 2 * kHalo;      // 160 rows
constexpr int kRowBytes = kC * 2;                      // 128 B: one position, 64 fp16 channels = one 128B-swizzle row
constexpr int kPlanes = kC / 8;                        // 8 sixteen-byte chunks per row
constexpr int kStageBytes = kRows * kRowBytes;         // 20480
constexpr int kStages = 2;
constexpr int kTapBytes = kC * kRowBytes;               // 8192: [cout 64][128 B]
constexpr int kWBytes = 9 * kTapBytes;                 // 73728
constexpr int kOutBytes = kBoards * kPos * kRowBytes;  // 16384: one output tile in the global board layout
constexpr int kBoardHalves = kC * kPos;                // 4096 fp16 per board
constexpr int kAccCols = 64;
constexpr int kThreads = 384;             // 4 control warps + 8 epilogue warps
constexpr int kEpiWarps = 8;

struct Smem {
    // offsets
    static constexpr int w = 0;
    static constexpr int a = 2 * kWBytes;                                  // two weight sets: layer l+1 is prefetched while layer l multiplies
    static constexpr int out = a + kStages * kStageBytes;                  // 2 output tiles (2 boards x 8 KB) staged for the bulk store
    static constexpr int bias = out + 2 * kOutBytes;                       // [kTowerMaxLayers][64] floats
    static constexpr int bars = bias + kTowerMaxLayers * kC * 4;           // 8-byte aligned
    static constexpr int tmem_ptr = bars + 64 * 8;
    static constexpr int total = tmem_ptr + 16;
};
static_assert(Smem::total <= 232448, "shared memory budget");

// kind::f16 with fp16 operands (format 0), fp32 accumulate, A and B K-major, M = 128, N = 64
// (cute::UMMA::InstrDescriptor)
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kAccCols >> 3) << 17) | ((128u >> 4) << 24);

MZ_DEVINL void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate) : "memory");
}
MZ_DEVINL void umma_f16_words(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(kIdesc), "r"(accumulate), "r"(kDescHi) : "memory");
}
}  // namespace

// activation buffers hold fp16 (the host side types them float*: 2048 float slots per board)
MZ_DEVINL const __half* tower_board(const TowerArgs& a, int buf, int g) {
    const __half* base = reinterpret_cast<const __half*>(a.buf[buf]);
    if (buf == 0 && a.gather_parent)
        return base + ((size_t)g * a.pool_stride + a.gather_parent[g]) * (size_t)kBoardHalves;
    return base + (size_t)g * kBoardHalves;
}

__global__ void __launch_bounds__(kThreads, 1) conv_tower_tc_kernel(const __grid_constant__ TowerArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_w = s_base + Smem::w, s_a = s_base + Smem::a;
    float* s_bias = reinterpret_cast<float*>(smem + Smem::bias);
    const uint32_t bars = s_base + Smem::bars;
    // weight set = layer & 1
    auto bar_w_full = [&](int set, int tap) { return bars + 8u * (set * 9 + tap); };          // weights of (layer, tap) landed
    auto bar_w_empty = [&](int set, int tap) { return bars + 8u * (18 + set * 9 + tap); };    // last MMA of the layer on this tap done
    auto bar_a_full = [&](int s) { return bars + 8u * (36 + s); };
    auto bar_a_empty = [&](int s) { return bars + 8u * (38 + s); };
    auto bar_acc_full = [&](int s) { return bars + 8u * (40 + s); };
    auto bar_acc_empty = [&](int s) { return bars + 8u * (42 + s); };
    auto bar_tile_done = [&](int k) { return bars + 8u * (44 + k); };      // layer output of my k-th tile stored
    auto bar_out_full = [&](int st) { return bars + 8u * (52 + st); };     // output tile staged in shared memory
    auto bar_out_empty = [&](int st) { return bars + 8u * (54 + st); };    // ... and drained by the bulk store
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Smem::tmem_ptr);

    const int n_tiles = (a.n + kBoards - 1) / kBoards;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int L = a.n_layers;

    // ---- one-time setup
    // zero the halo rows (the board rows are always overwritten by the bulk copies)
    for (int i = threadIdx.x; i < kStages * 2 * kHalo * (kRowBytes / 16); i += kThreads) {
        const int chunk = i % (kRowBytes / 16), r = (i / (kRowBytes / 16)) % (2 * kHalo), st = i / ((kRowBytes / 16) * 2 * kHalo);
        const int row = r < kHalo ? r : kRows - 2 * kHalo + r;
        reinterpret_cast<uint4*>(smem + Smem::a + st * kStageBytes + row * kRowBytes)[chunk] = make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < L * kC; i += kThreads) {
        const float* b = a.layer[i / kC].bias;
        s_bias[i] = b ? b[i % kC] : 0.0f;
    }
    if (threadIdx.x == 0) {
        for (int t = 0; t < 18; ++t) { mbar_init(bar_w_full(t / 9, t % 9), 1); mbar_init(bar_w_empty(t / 9, t % 9), 1); }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_a_full(s), 1);
            mbar_init(bar_a_empty(s), 1);
            mbar_init(bar_acc_full(s), 1);
            mbar_init(bar_acc_empty(s), kEpiWarps);   // one arrival per epilogue warp
        }
        for (int k = 0; k < kTowerMaxTiles; ++k) mbar_init(bar_tile_done(k), 1);
        for (int st = 0; st < 2; ++st) { mbar_init(bar_out_full(st), kEpiWarps); mbar_init(bar_out_empty(st), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic zero-fill -> async proxy readers
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(s_base + Smem::tmem_ptr), "r"(2 * kAccCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ================= producer =================
        int it = 0;
        // weights: two sets of nine tap slots; layer l uses set l & 1 and is loaded one layer ahead, as soon as the
        // last tile of layer l-2 has multiplied with the slot (so a layer never starts by waiting for its weights)
        auto load_weights = [&](int l) {
            const int set = l & 1, use = l >> 1;
            if (use > 0) mbar_wait(bar_w_empty(set, lane), (uint32_t)((use - 1) & 1));
            mbar_expect_tx(bar_w_full(set, lane), kTapBytes);
            bulk_g2s(s_w + set * kWBytes + lane * kTapBytes,
                     reinterpret_cast<const unsigned char*>(a.layer[l].w) + (size_t)lane * kTapBytes, kTapBytes, bar_w_full(set, lane));
        };
        if (my_tiles > 0 && lane < 9) { load_weights(0); if (L > 1) load_weights(1); }
        for (int l = 0; l < L; ++l) {
            if (l >= 1 && l + 1 < L && my_tiles > 0 && lane < 9) load_weights(l + 1);
            __syncwarp();
            const int in_buf = a.layer[l].in_buf;
            for (int k = 0; k < my_tiles; ++k, ++it) {
                const int tile = blockIdx.x + k * gridDim.x;
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                if (lane == 0) {
                    mbar_wait(bar_a_empty(s), ph ^ 1);
                    if (l > 0) {
                        // this tile's input was written by this CTA's epilogue in the previous layer
                        mbar_wait(bar_tile_done(k), (uint32_t)((l - 1) & 1));
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                }
                __syncwarp();
                const int nb = min(kBoards, a.n - tile * kBoards);
                if (a.debug_skip & 2) { if (lane == 0) mbar_arrive(bar_a_full(s)); __syncwarp(); continue; }
                if (lane == 0) mbar_expect_tx(bar_a_full(s), (uint32_t)nb * kPos * kRowBytes);
                __syncwarp();
                if (lane < nb) {                                           // one 8 KB bulk copy per board
                    const __half* src = tower_board(a, in_buf, tile * kBoards + lane);
                    bulk_g2s(s_a + s * kStageBytes + (kHalo + lane * kPos) * kRowBytes, src, kPos * kRowBytes, bar_a_full(s));
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        int it = 0;
        for (int l = 0; l < L; ++l) {
            for (int k = 0; k < my_tiles; ++k, ++it) {
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(bar_acc_empty(s), ph ^ 1);
                mbar_wait(bar_a_full(s), ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + (uint32_t)(s * kAccCols);
                    // Descriptors differ only in the 14-bit start-address field: build the constant words once and
                    // derive every MMA's descriptor with one add (the single issuing thread is latency-bound, so the
                    // instruction count per MMA is what sets the tensor-pipe duty cycle).
                    const uint32_t a16 = ((s_a + s * kStageBytes + kHalo * kRowBytes) >> 4) | kDescLoFlags;     // tile row 0, in 16-byte units
                    const uint32_t w16 = ((s_w + (uint32_t)((l & 1) * kWBytes)) >> 4) | kDescLoFlags;
                    uint32_t acc = 0;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (a.debug_skip & 1) break;
                        if (k == 0) { mbar_wait(bar_w_full(l & 1, tap), (uint32_t)((l >> 1) & 1)); tc_fence_after(); }
                        constexpr int kRow16 = kRowBytes / 16;                       // 8 sixteen-byte units per row
                        const int shift = (tap / 3 - 1) * 8 + (tap % 3 - 1);         // compile-time after unrolling
#pragma unroll
                        for (int ks = 0; ks < kC / 16; ++ks) {                       // K = 16 channels = 32 bytes of the row
                            const uint32_t alo = a16 + (uint32_t)(shift * kRow16 + ks * 2);      // < 2^14: never carries into the flag bits
                            const uint32_t blo = w16 + (uint32_t)(tap * (kTapBytes / 16) + ks * 2);
                            umma_f16_words(d, alo, blo, acc);
                            acc = 1;
                        }
                        if (k == my_tiles - 1) umma_commit(bar_w_empty(l & 1, tap));   // slot reusable by layer l + 2
                    }
                    if (a.debug_skip & 1) { if (k == my_tiles - 1) for (int tap = 0; tap < 9; ++tap) umma_commit(bar_w_empty(l & 1, tap)); }
                    umma_commit(bar_a_empty(s));          // smem stage reusable once the MMAs have read it
                    umma_commit(bar_acc_full(s));         // accumulator complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int q = warp & 3;                       // TMEM lane quarter
        const int half = (warp - 4) >> 2;             // accumulator columns [32*half, 32*half+32)
        const int row = q * 32 + lane;                // tile row = TMEM lane
        const int b = row / kPos, p = row % kPos;
        const int y = p / 8 - 1, x = p % 8;
        const bool inside = (y >= 0 && y < a.H && x < a.W);
        constexpr int kJ = kPlanes / 2;               // 16-byte chunks (8 channels) handled by this warp: 32 channels
        const size_t row_off = (size_t)p * kC;        // this position's 128-byte row, in fp16 elements
        const int sw = p & 7;                         // chunk c of the row is stored at chunk c ^ sw
        int it = 0;
        for (int l = 0; l < L; ++l) {
            const TowerLayer& ly = a.layer[l];
            const float* bias = s_bias + l * kC + half * 32;
            for (int k = 0; k < my_tiles; ++k, ++it) {
                const int tile = blockIdx.x + k * gridDim.x;
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                const int g = tile * kBoards + b;
                const bool live = inside && g < a.n;
                // ---- prefetch everything that does not depend on the accumulator
                uint4 res[kJ];                          // residual, 8 fp16 per channel group
                float act_scale = 0.0f;
#pragma unroll
                for (int j = 0; j < kJ; ++j) res[j] = make_uint4(0, 0, 0, 0);
                if (live) {
                    if (ly.res_buf >= 0 && !(a.debug_skip & 8)) {
                        const __half* rp = tower_board(a, ly.res_buf, g) + row_off;
#pragma unroll
                        for (int j = 0; j < kJ; ++j) res[j] = __ldcg(reinterpret_cast<const uint4*>(rp + (((half * kJ + j) ^ sw) << 3)));   // written by bulk stores: not through L1
                    }
                    if (ly.action_table) act_scale = __fdiv_rn((float)a.action[g], (float)a.A);
                }
                mbar_wait(bar_acc_full(s), ph);
                tc_fence_after();
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kAccCols + half * 32);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty(s));        // accumulator half read: may be overwritten
                // the output tile is staged in shared memory in the global board layout and leaves with one bulk store
                // per board (warp 3): the epilogue never waits for global memory
                const int st = it & 1;
                mbar_wait(bar_out_empty(st), ((uint32_t)(it >> 1) & 1u) ^ 1u);
                if (g < a.n) {
                    unsigned char* dst = smem + Smem::out + st * kOutBytes + (b * kPos + p) * kRowBytes;
                    const float* atab = ly.action_table ? ly.action_table + (size_t)p * kC + half * 32 : nullptr;
#pragma unroll
                    for (int j = 0; j < kJ; ++j) {
                        uint4 o = make_uint4(0, 0, 0, 0);
                        if (inside) {
                            float r[8];
                            const uint32_t rw[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 rf = unpack_f16x2(rw[e]);
                                r[2 * e + 0] = __uint_as_float(v[8 * j + 2 * e + 0]) + bias[8 * j + 2 * e + 0] + rf.x;
                                r[2 * e + 1] = __uint_as_float(v[8 * j + 2 * e + 1]) + bias[8 * j + 2 * e + 1] + rf.y;
                            }
                            if (atab) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) r[e] = fmaf(act_scale, atab[8 * j + e], r[e]);
                            }
                            if (ly.relu) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) r[e] = fmaxf(r[e], 0.0f);
                            }
                            o = make_uint4(pack_f16x2(r[0], r[1]), pack_f16x2(r[2], r[3]), pack_f16x2(r[4], r[5]), pack_f16x2(r[6], r[7]));
                        }
                        *reinterpret_cast<uint4*>(dst + (((half * kJ + j) ^ sw) << 4)) = o;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic smem writes -> bulk-copy reader
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_out_full(st));
            }
        }
    } else if (warp == 3) {
        // ================= output store =================
        if (lane == 0) {
            int it = 0;
            for (int l = 0; l < L; ++l) {
                __half* out = reinterpret_cast<__half*>(a.buf[a.layer[l].out_buf]);
                for (int k = 0; k < my_tiles; ++k, ++it) {
                    const int st = it & 1;
                    const int tile = blockIdx.x + k * gridDim.x;
                    const int nb = min(kBoards, a.n - tile * kBoards);
                    mbar_wait(bar_out_full(st), (uint32_t)(it >> 1) & 1u);
                    if (!(a.debug_skip & 4)) {
                        for (int b = 0; b < nb; ++b)
                            bulk_s2g(out + (size_t)(tile * kBoards + b) * kBoardHalves,
                                     s_base + Smem::out + st * kOutBytes + b * (kPos * kRowBytes), kPos * kRowBytes);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    // the staging buffer is free as soon as the copy engine has READ it; the tile counts as stored
                    // (next layer may load it) only when the writes are complete - tracked one store behind, so two
                    // stores are in flight inside a layer, and drained at the end of every layer
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    mbar_arrive(bar_out_empty(st));
                    if (k > 0) {
                        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                        if (l + 1 < L) { __threadfence(); mbar_arrive(bar_tile_done(k - 1)); }
                    }
                    if (k == my_tiles - 1) {
                        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                        if (l + 1 < L) { __threadfence(); mbar_arrive(bar_tile_done(k)); }
                    }
                }
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kAccCols) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// Resident tower: the same convolutions, but the activations of a CTA's tiles never leave shared memory between the
// layers.  Two activation buffers B0 / B1 hold the CTA's (up to kResTiles) tiles back to back in the board layout
// (consecutive boards are separated by their own zero padding rows, so a tap window that leaves a board only reads
// zeros); layer l reads B[l & 1] and writes B[(l & 1) ^ 1]; the second conv of a block adds the residual IN PLACE (the
// block input is what the output buffer still holds, and a row is read and overwritten by the same thread).  Global
// memory is touched three times per tower: bulk loads of the input boards, the weight taps (one slot set, refilled
// for layer l+1 while the last tile of layer l multiplies), bulk stores of the last layer's tiles.  There is no
// activation ring, no per-tile store fence and no "tile stored" handshake: MMA(l+1, k) only waits for the epilogue
// of (l, k); write-after-read hazards are excluded by the in-order completion of the MMAs (the epilogue of layer
// l+2 starts after a tcgen05.commit that follows every MMA of layer l+1).
// ------------------------------------------------------------------------------------------------------------------
constexpr int kResTiles = 4;
constexpr int kResRows = kResTiles * kBoards * kPos + 2 * kHalo;      // 544
constexpr int kResBufBytes = kResRows * kRowBytes;                    // 69632 (multiple of 1024: swizzle phase kept)

struct SmemR {
    static constexpr int w = 0;
    static constexpr int act = kWBytes;                                    // B0 | B1
    static constexpr int bias = act + 2 * kResBufBytes;
    static constexpr int bars = bias + kTowerMaxLayers * kC * 4;
    static constexpr int tmem_ptr = bars + 48 * 8;
    static constexpr int total = tmem_ptr + 16;
};
static_assert(SmemR::total <= 232448, "shared memory budget");
static_assert(SmemR::act % 1024 == 0 && kResBufBytes % 1024 == 0, "activation buffers must keep the 1024-byte swizzle phase");

__global__ void __launch_bounds__(kThreads, 1) conv_tower_resident_kernel(const __grid_constant__ TowerArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_w = s_base + SmemR::w, s_act = s_base + SmemR::act;
    float* s_bias = reinterpret_cast<float*>(smem + SmemR::bias);
    const uint32_t bars = s_base + SmemR::bars;
    auto bar_w_full = [&](int tap) { return bars + 8u * tap; };
    auto bar_w_empty = [&](int tap) { return bars + 8u * (9 + tap); };
    auto bar_in_full = [&](int k) { return bars + 8u * (18 + k); };         // input boards of my k-th tile landed
    auto bar_tile_ready = [&](int k) { return bars + 8u * (22 + k); };      // epilogue of (layer, k) wrote its rows (one phase per layer)
    auto bar_out_ready = [&](int k) { return bars + 8u * (26 + k); };       // last layer's rows of tile k written
    auto bar_acc_full = [&](int s) { return bars + 8u * (30 + s); };
    auto bar_acc_empty = [&](int s) { return bars + 8u * (32 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + SmemR::tmem_ptr);

    const int n_tiles = (a.n + kBoards - 1) / kBoards;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int L = a.n_layers;

    // ---- one-time setup: zero halos of both buffers, biases, barriers, TMEM
    for (int i = threadIdx.x; i < 2 * 2 * kHalo * (kRowBytes / 16); i += kThreads) {
        const int chunk = i % (kRowBytes / 16), r = (i / (kRowBytes / 16)) % (2 * kHalo), bf = i / ((kRowBytes / 16) * 2 * kHalo);
        const int row = r < kHalo ? r : kResRows - 2 * kHalo + r;
        reinterpret_cast<uint4*>(smem + SmemR::act + bf * kResBufBytes + row * kRowBytes)[chunk] = make_uint4(0, 0, 0, 0);
    }
    for (int i = threadIdx.x; i < L * kC; i += kThreads) {
        const float* b = a.layer[i / kC].bias;
        s_bias[i] = b ? b[i % kC] : 0.0f;
    }
    if (threadIdx.x == 0) {
        for (int t = 0; t < 9; ++t) { mbar_init(bar_w_full(t), 1); mbar_init(bar_w_empty(t), 1); }
        for (int k = 0; k < kResTiles; ++k) {
            mbar_init(bar_in_full(k), 1);
            mbar_init(bar_tile_ready(k), kEpiWarps / 2);
            mbar_init(bar_out_ready(k), kEpiWarps / 2);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_acc_full(s), 1); mbar_init(bar_acc_empty(s), kEpiWarps / 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(s_base + SmemR::tmem_ptr), "r"(2 * kAccCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= producer: weights of every layer, input boards once =================
        pdl_launch_dependents();
        if (my_tiles > 0 && lane < 9) {
            mbar_expect_tx(bar_w_full(lane), kTapBytes);
            bulk_g2s(s_w + lane * kTapBytes, reinterpret_cast<const unsigned char*>(a.layer[0].w) + (size_t)lane * kTapBytes, kTapBytes,
                     bar_w_full(lane));
        }
        pdl_wait();                                    // weights are constants; the boards come from the previous kernel
        __syncwarp();
        const int in_buf = a.layer[0].in_buf;
        for (int k = 0; k < my_tiles; ++k) {
            const int tile = blockIdx.x + k * gridDim.x;
            const int nb = min(kBoards, a.n - tile * kBoards);
            if (lane == 0) mbar_expect_tx(bar_in_full(k), (uint32_t)nb * kPos * kRowBytes);
            __syncwarp();
            if (lane < nb)
                bulk_g2s(s_act + (kHalo + (k * kBoards + lane) * kPos) * kRowBytes, tower_board(a, in_buf, tile * kBoards + lane),
                         kPos * kRowBytes, bar_in_full(k));
        }
        for (int l = 1; l < L; ++l) {
            if (my_tiles > 0 && lane < 9) {
                mbar_wait(bar_w_empty(lane), (uint32_t)((l - 1) & 1));        // last tile of layer l-1 is done with this tap
                mbar_expect_tx(bar_w_full(lane), kTapBytes);
                bulk_g2s(s_w + lane * kTapBytes, reinterpret_cast<const unsigned char*>(a.layer[l].w) + (size_t)lane * kTapBytes,
                         kTapBytes, bar_w_full(lane));
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        int it = 0;
        for (int l = 0; l < L; ++l) {
            for (int k = 0; k < my_tiles; ++k, ++it) {
                const int s = it & 1;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                if (l == 0) mbar_wait(bar_in_full(k), 0);
                else mbar_wait(bar_tile_ready(k), (uint32_t)((l - 1) & 1));
                mbar_wait(bar_acc_empty(s), ph ^ 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + (uint32_t)(s * kAccCols);
                    const uint32_t a16 = ((s_act + (uint32_t)((l & 1) * kResBufBytes + (kHalo + k * kBoards * kPos) * kRowBytes)) >> 4) | kDescLoFlags;
                    const uint32_t w16 = (s_w >> 4) | kDescLoFlags;
                    uint32_t acc = 0;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (a.debug_skip & 1) break;
                        if (k == 0) { mbar_wait(bar_w_full(tap), (uint32_t)(l & 1)); tc_fence_after(); }
                        constexpr int kRow16 = kRowBytes / 16;
                        const int shift = (tap / 3 - 1) * 8 + (tap % 3 - 1);
#pragma unroll
                        for (int ks = 0; ks < kC / 16; ++ks) {
                            const uint32_t alo = a16 + (uint32_t)(shift * kRow16 + ks * 2);      // < 2^14: never carries into the flag bits
                            const uint32_t blo = w16 + (uint32_t)(tap * (kTapBytes / 16) + ks * 2);
                            umma_f16_words(d, alo, blo, acc);
                            acc = 1;
                        }
                        if (k == my_tiles - 1) umma_commit(bar_w_empty(tap));
                    }
                    if (a.debug_skip & 1) { if (k == my_tiles - 1) for (int tap = 0; tap < 9; ++tap) umma_commit(bar_w_empty(tap)); }
                    umma_commit(bar_acc_full(s));
                }
                __syncwarp();
            }
        }
    } else if (warp == 3) {
        // ================= output store: the last layer's tiles leave as bulk copies =================
        if (lane == 0) {
            __half* out = reinterpret_cast<__half*>(a.buf[a.layer[L - 1].out_buf]);
            const uint32_t src = s_act + (uint32_t)((((L - 1) & 1) ^ 1) * kResBufBytes + kHalo * kRowBytes);
            for (int k = 0; k < my_tiles; ++k) {
                const int tile = blockIdx.x + k * gridDim.x;
                const int nb = min(kBoards, a.n - tile * kBoards);
                mbar_wait(bar_out_ready(k), 0);
                if (!(a.debug_skip & 4))
                    for (int b = 0; b < nb; ++b)
                        bulk_s2g(out + (size_t)(tile * kBoards + b) * kBoardHalves, src + (uint32_t)((k * kBoards + b) * kPos * kRowBytes),
                                 kPos * kRowBytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp >= 4) {
        // ================= epilogue: two groups of four warps, group s owns accumulator stage s =================
        pdl_wait();                                    // reads a.action (written by the tree kernel)
        // (consecutive tiles are drained by different groups, so the wait -> tcgen05.ld -> store -> fence -> arrive
        // chain of one tile overlaps the next tile's)
        const int q = warp & 3;                       // TMEM lane quarter
        const int grp = (warp - 4) >> 2;              // tiles with (it & 1) == grp
        const int row = q * 32 + lane;                // tile row = TMEM lane
        const int b = row / kPos, p = row % kPos;
        const int y = p / 8 - 1, x = p % 8;
        const bool inside = (y >= 0 && y < a.H && x < a.W);
        const int sw = p & 7;                         // chunk c of the row is stored at chunk c ^ sw
        int it = 0;
        for (int l = 0; l < L; ++l) {
            const TowerLayer& ly = a.layer[l];
            const float* bias = s_bias + l * kC;
            const bool last = l == L - 1;
            for (int k = 0; k < my_tiles; ++k, ++it) {
                if ((it & 1) != grp) continue;
                const int tile = blockIdx.x + k * gridDim.x;
                const int s = grp;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                const int g = tile * kBoards + b;
                const bool live = inside && g < a.n;
                // this thread's row in the output buffer (residual source and destination)
                unsigned char* orow = smem + SmemR::act + ((l & 1) ^ 1) * kResBufBytes + (kHalo + k * kBoards * kPos + row) * kRowBytes;
                float act_scale = 0.0f;
                if (live && ly.action_table) act_scale = __fdiv_rn((float)a.action[g], (float)a.A);
                mbar_wait(bar_acc_full(s), ph);
                tc_fence_after();
                uint32_t v[64];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * kAccCols);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t* w = v + 32 * h;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                          "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]),
                          "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
                          "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
                        : "r"(taddr + (uint32_t)(32 * h)));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty(s));        // accumulator rows read: may be overwritten
                const float* atab = ly.action_table ? ly.action_table + (size_t)p * kC : nullptr;
#pragma unroll
                for (int j = 0; j < kPlanes; ++j) {
                    uint4* slot = reinterpret_cast<uint4*>(orow + ((j ^ sw) << 4));
                    uint4 o = make_uint4(0, 0, 0, 0);
                    if (live) {
                        uint4 res = make_uint4(0, 0, 0, 0);
                        if (ly.res_buf >= 0) res = *slot;               // block input, added in place
                        float r[8];
                        const uint32_t rw[4] = {res.x, res.y, res.z, res.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 rf = unpack_f16x2(rw[e]);
                            r[2 * e + 0] = __uint_as_float(v[8 * j + 2 * e + 0]) + bias[8 * j + 2 * e + 0] + rf.x;
                            r[2 * e + 1] = __uint_as_float(v[8 * j + 2 * e + 1]) + bias[8 * j + 2 * e + 1] + rf.y;
                        }
                        if (atab) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) r[e] = fmaf(act_scale, atab[8 * j + e], r[e]);
                        }
                        if (ly.relu) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) r[e] = fmaxf(r[e], 0.0f);
                        }
                        o = make_uint4(pack_f16x2(r[0], r[1]), pack_f16x2(r[2], r[3]), pack_f16x2(r[4], r[5]), pack_f16x2(r[6], r[7]));
                    }
                    *slot = o;                                         // zeros on padding rows and missing boards
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic smem writes -> tcgen05 / bulk-copy readers
                __syncwarp();
                if (lane == 0) mbar_arrive(last ? bar_out_ready(k) : bar_tile_ready(k));
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kAccCols) : "memory");
}

// the resident kernel needs the block structure: every residual is the input of the layer before, buffers alternate
static bool tower_is_resident_shape(const TowerArgs& a, int tiles_per_cta) {
    if (a.n_layers < 2 || tiles_per_cta > kResTiles) return false;
    for (int l = 0; l < a.n_layers; ++l) {
        const TowerLayer& t = a.layer[l];
        if (l > 0 && t.in_buf != a.layer[l - 1].out_buf) return false;
        if (t.res_buf >= 0 && (l == 0 || t.res_buf != a.layer[l - 1].in_buf)) return false;
    }
    return true;
}

static int tower_ctas_per_sm() {
    static int per_sm = 0;
    if (per_sm == 0) {
        int occ = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv_tower_tc_kernel, kThreads, Smem::total) != cudaSuccess || occ < 1) occ = 1;
        if (getenv("MZ_TC_VERBOSE")) fprintf(stderr, "[conv_tower_tc] occupancy %d CTAs/SM, %d B shared per CTA\n", occ, Smem::total);
        const char* e = getenv("MZ_TC_CTAS");           // A/B switch: force one CTA per SM
        if (e && atoi(e) >= 1 && atoi(e) < occ) occ = atoi(e);
        per_sm = occ > 2 ? 2 : occ;
    }
    return per_sm;
}

cudaError_t launch_conv_tower_tc(const TowerArgs& a, int sm_count, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tower_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::total);
        if (e != cudaSuccess) return e;
        // all of the SM's unified L1/shared array as shared memory, so two CTAs fit side by side
        e = cudaFuncSetAttribute(conv_tower_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (a.n_layers < 1 || a.n_layers > kTowerMaxLayers) return cudaErrorInvalidValue;
    const int n_tiles = (a.n + kBoards - 1) / kBoards;
    // two co-resident CTAs per SM (fp16 operands leave room): their MMA streams interleave on the tensor core
    const int slots = sm_count * tower_ctas_per_sm();
    const int grid = n_tiles < slots ? n_tiles : slots;
    if (a.n_layers > 1 && (n_tiles + grid - 1) / grid > kTowerMaxTiles) return cudaErrorInvalidConfiguration;
    const char* nr = getenv("MZ_TC_NO_RESIDENT");            // A/B switch (tests compare both kernels)
    const bool no_resident = nr && nr[0] == '1';
    if (!no_resident && tower_is_resident_shape(a, (n_tiles + sm_count - 1) / sm_count)) {
        static bool attr_r = false;
        if (!attr_r) {
            cudaError_t e = cudaFuncSetAttribute(conv_tower_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemR::total);
            if (e != cudaSuccess) return e;
            attr_r = true;
        }
        const int grid_r = n_tiles < sm_count ? n_tiles : sm_count;
        cudaError_t e = launch_chained(conv_tower_resident_kernel, dim3(grid_r), dim3(kThreads), SmemR::total, stream, a);
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    cudaError_t e = launch_chained(conv_tower_tc_kernel, dim3(grid), dim3(kThreads), Smem::total, stream, a);
    return e != cudaSuccess ? e : cudaGetLastError();
}

int conv_tc_max_boards_fused(int sm_count) { return sm_count * kTowerMaxTiles * kBoards; }   // conservative: one CTA per SM

bool conv_tc_supported(int C, int H, int W) { return C == kC && H >= 1 && H <= 6 && W >= 1 && W <= 7; }
int conv_tc_board_elems(bool split) { return split ? kBoardHalves : kBoardHalves / 2; }      // float slots per board (fp16 plane, or x_h | x_l planes)

}  // namespace mz
