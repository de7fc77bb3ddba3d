// Device helpers shared by the tcgen05 tower kernels (conv_tc.cu, conv_x3.cu): mbarriers, bulk copies (TMA unit),
// leader election, tcgen05 fences / commit, shared-memory matrix descriptor words.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace mz {
namespace tc {

MZ_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

MZ_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
MZ_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
MZ_DEVINL void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
MZ_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
MZ_DEVINL void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// exactly one lane of a converged warp (ptxas then knows the tcgen05 operands come from a single thread and
// moves them to uniform registers without a broadcast loop)
MZ_DEVINL bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
MZ_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MZ_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// descriptor words shared by every A / B descriptor of this kernel (K-major SWIZZLE_128B, SBO = 1024 B, version 1);
// the hardware applies the 128B swizzle on absolute shared-memory address bits, so row-shifted tap windows need
// no base_offset (checked: tests/test_conv_gpu.py is exact with base_offset = 0 and wrong with the row phase)
constexpr uint32_t kDescLoFlags = 1u << 16;                                           // LBO field = 1 (unused)
constexpr uint32_t kDescHi = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);      // SBO | version | SWIZZLE_128B

// shared -> global bulk copy (TMA unit), tracked by the thread's bulk async-group
MZ_DEVINL void bulk_s2g(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
MZ_DEVINL void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// two fp32 -> packed fp16x2, round to nearest even, saturating to the finite range
MZ_DEVINL uint32_t pack_f16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
MZ_DEVINL float2 unpack_f16x2(uint32_t v) {
    __half2 h = *reinterpret_cast<__half2*>(&v);
    return __half22float2(h);
}


}  // namespace tc
}  // namespace mz
