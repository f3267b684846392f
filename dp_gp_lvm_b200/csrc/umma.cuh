// tcgen05 / TMEM primitives (sm_100a inline PTX) for the int8 sliced reductions of the fused psi2 backward.
// SASS: tcgen05.mma -> UTCIMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS.
//
// Shared-memory matrix descriptors, no swizzle ("interleave"): the operand is a grid of 128-byte core blocks.
//   K-major  operand: block = 8 rows (M or N index) x 16 bytes of K, the 8 rows 16 bytes apart;
//                     SBO = byte distance between 8-row groups, LBO = byte distance between 16-byte K chunks.
//   MN-major operand: block = 8 K indices x 16 bytes of M / N, the 8 K indices 16 bytes apart;
//                     SBO = byte distance between 16-element M / N groups, LBO = byte distance between 8-index K groups.
// One int8 MMA consumes K = 32.  Pinned on the hardware by csrc/microbench/umma_i8_probe.cu.
#pragma once
#include "common.cuh"

namespace dpgp {

__device__ __forceinline__ bool elect_one() {
  unsigned pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Warp-collective.  ncols: power of two >= 32.  The base address (lane 0, first column) lands in *smem_dst.
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, int ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t base, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}

__device__ __forceinline__ uint64_t umma_smem_desc(const void* smem, unsigned lbo_bytes, unsigned sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  return d;                                        // base offset 0, LBO mode 0, layout type 0 = no swizzle
}
// kind::i8 instruction descriptor: int32 accumulator, u8 / s8 operands, K-major or MN-major operands
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n, bool a_signed, bool b_signed, bool a_mn_major, bool b_mn_major) {
  return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((a_mn_major ? 1u : 0u) << 15) |
         ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[smem] B[smem]; one thread issues
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// the mbarrier receives one arrival when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: lane (32 (warp % 4) + laneid), 16 / 32 consecutive 32-bit columns from taddr = (lane << 16) | column
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

}  // namespace dpgp
