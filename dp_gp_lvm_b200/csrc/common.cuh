// Shared definitions for the DP-GP-LVM CUDA kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "fast_exp.cuh"

namespace dpgp {

constexpr double kJitter = 1.0e-8;          // src/utils/constants.py:96
constexpr double kRClamp = -3.0e8;          // keeps |E| < 2^31 ln2 for the exp argument reduction
constexpr int kMaxQ = 16;
constexpr int kMaxM = 256;

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// exp variants: 1 libdevice, 2 poly11 (fast_exp.cuh), 3 shuffle table
template <int EXPV>
struct Exp {
  ExpTable tab;
  __device__ __forceinline__ void init() { if (EXPV == 3) tab.init(); }
  // acc + exp(x)
  __device__ __forceinline__ double acc(double x, double a) const {
    if (EXPV == 1) return a + exp(x);
    if (EXPV == 2) return exp_acc(x, 1.0, a);
    return tab.exp_acc(x, a);
  }
  // w * exp(x)
  __device__ __forceinline__ double scaled(double x, double w) const {
    if (EXPV == 1) return w * exp(x);
    if (EXPV == 2) { int k; double p = exp_reduced(x, k); return p * (pow2i(k) * w); }
    double s; double p = tab.reduced(x, s); return p * (s * w);
  }
  __device__ __forceinline__ double value(double x) const {
    if (EXPV == 1) return exp(x);
    if (EXPV == 2) return exp_fast(x);
    return tab.exp(x);
  }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Upper-triangular enumeration of the Mt x Mt grid of 2x2 tiles: t -> (I, J), I <= J, row-major.
__host__ __device__ inline void tile_from_index(int t, int mt, int& ti, int& tj) {
  int i = 0, rem = t;
  while (rem >= mt - i) { rem -= mt - i; ++i; }
  ti = i; tj = i + rem;
}
__host__ __device__ inline int tile_index(int ti, int tj, int mt) { return ti * mt - ti * (ti - 1) / 2 + (tj - ti); }

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
