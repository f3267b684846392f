// Shared definitions for the DP-GP-LVM CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "fast_exp.cuh"

namespace dpgp {

constexpr double kJitter = 1.0e-8;          // src/utils/constants.py:96
constexpr double kRClamp = -3.0e8;          // keeps the exp argument inside the domain of the argument reduction (fast_exp.cuh)
constexpr int kMaxQ = 32;
constexpr int kMaxM = 256;

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// exp variants: 1 libdevice, 2 poly11 (fast_exp.cuh), 3 shuffle table, 4/5/6 shared-memory table (256/64/32 entries)
template <int EXPV>
struct Exp {
  ExpTable tab;
  const double* stab = nullptr;               // EXPV >= 4: 2^(j/S) in shared memory
  __device__ __forceinline__ void init(const double* smem_tab = nullptr) { if (EXPV == 3) tab.init(); stab = smem_tab; }
  // acc + exp(x)
  __device__ __forceinline__ double acc(double x, double a) const {
    if (EXPV == 1) return a + exp(x);
    if (EXPV == 2) return exp_acc(x, 1.0, a);
    if (EXPV >= 4) { const double xx[1] = {x}; double aa[1] = {a}; exp_tab_acc_k<exp_tab_bits(EXPV), 1>(stab, xx, aa); return aa[0]; }
    return tab.exp_acc(x, a);
  }
  // w * exp(x)
  __device__ __forceinline__ double scaled(double x, double w) const {
    if (EXPV == 1) return w * exp(x);
    if (EXPV == 2) { int k; double p = exp_reduced(x, k); return p * (pow2i(k) * w); }
    if (EXPV >= 4) { const double xx[1] = {x}, ww[1] = {w}; double oo[1]; exp_tab_scaled_k<exp_tab_bits(EXPV), 1>(stab, xx, ww, oo); return oo[0]; }
    double s; double p = tab.reduced(x, s); return p * (s * w);
  }
  __device__ __forceinline__ double value(double x) const {
    if (EXPV == 1) return exp(x);
    if (EXPV == 2) return exp_fast(x);
    if (EXPV >= 4) return scaled(x, 1.0);
    return tab.exp(x);
  }
};

// Copies the host-built 2^(j/S) table (handle workspace) into shared memory; caller synchronises.
__device__ __forceinline__ void load_exp_table(double* smem_tab, const double* __restrict__ gtab) {
  for (int i = threadIdx.x; i < kExpTabSize; i += blockDim.x) smem_tab[i] = gtab[i];
}

// K independent "w * exp(x)" evaluated in lockstep: every Horner step is issued for all K chains before the
// next one, so the FP64 pipe always sees K independent DFMAs (DFMA latency is 8 cycles = 4 issue slots; ptxas
// does not interleave separately inlined exp bodies by itself -- profiles/r01_psi2.md).
template <int EXPV, int K>
__device__ __forceinline__ void exp_scaled_k(const Exp<EXPV>& ex, const double (&x)[K], const double (&w)[K], double (&out)[K]) {
  if (EXPV == 1) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = w[k] * exp(x[k]);
    return;
  }
  if (EXPV == 3) {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = ex.scaled(x[k], w[k]);
    return;
  }
  if (EXPV >= 4) { exp_tab_scaled_k<exp_tab_bits(EXPV), K>(ex.stab, x, w, out); return; }
  const double MAGIC = 6755399441055744.0, L2E = 0x1.71547652b82fep+0, NLN2 = -0x1.62e42fefa39efp-1;
  double t[K], r[K], q[K]; int kk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(x[k], L2E, MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) { kk[k] = __double2loint(t[k]); t[k] -= MAGIC; }
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(t[k], NLN2, x[k]);
  const double c[10] = {0x1.af389ecfc4b9cp-26, 0x1.28917c89a43a7p-22, 0x1.71de0db2f6b19p-19, 0x1.a019b9149a41cp-16,
                        0x1.a01a01a7c2efep-13, 0x1.6c16c17889ef1p-10, 0x1.11111111109b5p-7, 0x1.5555555553d68p-5,
                        0x1.5555555555556p-3, 0x1.0000000000001p-1};
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(c[0], r[k], c[1]);
#pragma unroll
  for (int j = 2; j < 10; ++j) {
#pragma unroll
    for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], c[j]);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = q[k] * (pow2i(kk[k]) * w[k]);
}

// acc[k] += exp(x[k]) for K chains (lockstep for the table variant).
template <int EXPV, int K>
__device__ __forceinline__ void exp_acc_k(const Exp<EXPV>& ex, const double (&x)[K], double (&acc)[K]) {
  if constexpr (EXPV >= 4) {
    exp_tab_acc_k<exp_tab_bits(EXPV), K>(ex.stab, x, acc);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = ex.acc(x[k], acc[k]);
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS UBLKCP / SYNCS) --------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`; 16-byte aligned, size multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Upper-triangular enumeration of the Mt x Mt grid of 2x2 tiles: t -> (I, J), I <= J, row-major.
__host__ __device__ inline void tile_from_index(int t, int mt, int& ti, int& tj) {
  int i = 0, rem = t;
  while (rem >= mt - i) { rem -= mt - i; ++i; }
  ti = i; tj = i + rem;
}
__host__ __device__ inline int tile_index(int ti, int tj, int mt) { return ti * mt - ti * (ti - 1) / 2 + (tj - ti); }

// CTA c of the producing kernel owns items [items c / grid, items (c+1) / grid) of the (cluster, chunk) list, so the slots
// that can hold cluster b belong to a short, computable range of CTAs (the first version walked all grid * nseg slots per
// output: 33 us at the reference's small shapes).  Fixed order -> bitwise reproducible.
__host__ __device__ inline void cta_range_of_cluster(int b, int64_t nchunks, int b_count, int grid, int& c_lo, int& c_hi) {
  const int64_t items = nchunks * b_count;
  c_lo = (int)(((int64_t)b * nchunks * grid) / items) - 1;
  c_hi = (int)((((int64_t)b + 1) * nchunks * grid + items - 1) / items) + 1;
  if (c_lo < 0) c_lo = 0;
  if (c_hi > grid - 1) c_hi = grid - 1;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    double x = (l < nw) ? red[l] : 0.0;
    x = warp_sum(x);
    return x;
  }
  return 0.0;
}

}  // namespace dpgp
