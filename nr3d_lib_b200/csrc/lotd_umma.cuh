// lotd_umma.cuh -- the few tcgen05 / TMEM / mbarrier pieces shared by the fused encoder + decoder kernels (lotd_fused.cu forward,
// lotd_fused_bwd.cu backward): shared-memory matrix descriptors for the un-swizzled canonical layouts, the kind::f16 instruction
// descriptor, MMA issue / commit, bounded mbarrier waits and TMEM loads.  Layout facts (CUTLASS cute/atom/mma_traits_sm100.hpp:165-203
// restated): a "core matrix" is 8 rows x 16 bytes; with no swizzle
//   K-major  operand (rows = M or N, 16 bytes = 8 consecutive k):   LBO = byte step between the two k-halves of one MMA (k-blocks of 8),
//                                                                    SBO = byte step between 8-row groups along M / N;
//   MN-major operand (rows = k, 16 bytes = 8 consecutive m or n):    LBO = byte step between k-groups of 8, SBO = byte step between
//                                                                    groups of 8 along M / N.
// A tile stored as [row r][channel c] at (c / 8) * CB + (r / 8) * 128 + (r % 8) * 16 + (c % 8) * 2 is therefore BOTH a K-major operand
// with rows r (LBO = CB, SBO = 128) and an MN-major operand whose M / N index is c and whose K index is r (LBO = 128, SBO = CB) -- which
// is how the backward contracts over the points of a tile (weight gradients) without a transposed copy.
#pragma once
#include "lotd_pair.cuh"
#include <cuda_bf16.h>

namespace nr3d {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row, k) lives at  (k / 8) * LBO + (row / 8) * SBO + (row % 8) * 16 + (k % 8) * 2  bytes
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::f16: D = f32, A = B = bf16, shape M x N; a_mn / b_mn = 1: that operand is MN-major (else K-major)
__device__ __forceinline__ constexpr uint32_t umma_idesc(uint32_t M, uint32_t N, uint32_t a_mn = 0, uint32_t b_mn = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();  // never hang the device: a lost arrival becomes a launch error
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

#define NR3D_TMEM_LD16(taddr, v)                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                     \
                 : "r"(taddr))

#define NR3D_TMEM_LD8(taddr, v)                                                                  \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"        \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr))

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
}

}  // namespace nr3d
