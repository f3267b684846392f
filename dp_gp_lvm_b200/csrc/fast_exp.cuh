// FP64 exp for the psi-statistic kernels (sm_100a).
//
// B200 has no FP64 SFU: exp is software on the DFMA pipe, and it is the largest single cost of a psi2
// unit (one exp per (cluster, n, m<=m')).  libdevice exp() spends ~25 FP64-pipe instructions plus range
// checks; the variants here spend 14 (polynomial) or 10 (shuffle table) and fold the final 2^k scaling
// into the FMA that accumulates the result, so "acc += w * exp(x)" costs 15 / 11 FP64 issues.
//
// Accuracy (coefficients fitted with mpmath, see tools/fit_exp.py; measured on the GPU by
// csrc/microbench/fp64_peaks.cu against libdevice): approximation error 1.6e-17 (poly11) / 2.2e-19
// (table) relative, plus the argument-reduction error |k| * 2.3e-17 from using a single-word ln2
// (|k| <= 58 for every term larger than 1e-17 of the largest possible one), i.e. <= ~2e-15 relative.
//
// Domain: |x| < 1.4e9 (k must fit in 32 bits), x < 709.  Results below 2^-1022 are flushed to exactly 0.
// NaN / inf arguments are not supported (the psi kernels never produce them for finite inputs).
#pragma once
#include <cuda_runtime.h>

namespace dpgp {

// 2^k as a double, k clamped to the normal range; k <= -1023 gives exactly 0.0.
__device__ __forceinline__ double pow2i(int k) {
  int e = k + 1023;
  e = max(e, 0);
  e = min(e, 2046);
  return __hiloint2double(e << 20, 0);
}

// Argument reduction x = k ln2 + r, |r| <= ln2/2 ; returns p(r) ~ exp(r) and k.
__device__ __forceinline__ double exp_reduced(double x, int& k) {
  const double MAGIC = 6755399441055744.0;          // 1.5 * 2^52
  const double L2E = 0x1.71547652b82fep+0;
  const double NLN2 = -0x1.62e42fefa39efp-1;
  double t = fma(x, L2E, MAGIC);
  k = __double2loint(t);
  double kf = t - MAGIC;
  double r = fma(kf, NLN2, x);
  double q = 0x1.af389ecfc4b9cp-26;
  q = fma(q, r, 0x1.28917c89a43a7p-22);
  q = fma(q, r, 0x1.71de0db2f6b19p-19);
  q = fma(q, r, 0x1.a019b9149a41cp-16);
  q = fma(q, r, 0x1.a01a01a7c2efep-13);
  q = fma(q, r, 0x1.6c16c17889ef1p-10);
  q = fma(q, r, 0x1.11111111109b5p-7);
  q = fma(q, r, 0x1.5555555553d68p-5);
  q = fma(q, r, 0x1.5555555555556p-3);
  q = fma(q, r, 0x1.0000000000001p-1);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  return q;
}

// exp(x)
__device__ __forceinline__ double exp_fast(double x) {
  int k; double p = exp_reduced(x, k);
  return p * pow2i(k);
}

// acc + w * exp(x) for w == 1 (the multiply by w is only issued when w is not the literal 1.0)
__device__ __forceinline__ double exp_acc(double x, double w, double acc) {
  int k; double p = exp_reduced(x, k);
  double s = pow2i(k);
  if (w != 1.0) s *= w;
  return fma(p, s, acc);
}

// Shuffle-table variant: x = (32 e + j) ln2/32 + r, |r| <= ln2/64, exp(x) = 2^e * T[j] * p6(r).
// T lives one entry per lane (32 lanes = 32 entries) and is fetched with a warp shuffle, which has no
// bank conflicts.  All 32 lanes of the warp must be converged at every call.
struct ExpTable {
  double t;   // 2^(lane/32)
  __device__ __forceinline__ void init() { t = exp2((double)(threadIdx.x & 31) * (1.0 / 32.0)); }

  __device__ __forceinline__ double reduced(double x, double& s) const {
    const double MAGIC = 6755399441055744.0;
    const double L2E32 = 0x1.71547652b82fep+5;
    const double NLN2_32 = -0x1.62e42fefa39efp-6;
    double tt = fma(x, L2E32, MAGIC);
    int ki = __double2loint(tt);
    double kf = tt - MAGIC;
    double r = fma(kf, NLN2_32, x);
    double tj = __shfl_sync(0xffffffffu, t, ki & 31);
    int e = ki >> 5;                                   // floor division, matches j = ki & 31
    int hi = __double2hiint(tj) + (e << 20);           // tj in [1,2): exponent field 1023
    hi = (e < -1022) ? 0 : hi;
    int lo = (e < -1022) ? 0 : __double2loint(tj);
    s = __hiloint2double(hi, lo);
    double q = 0x1.6c16ffe57d9c9p-10;
    q = fma(q, r, 0x1.11114f8a7941cp-7);
    q = fma(q, r, 0x1.555555555194dp-5);
    q = fma(q, r, 0x1.555555554dd45p-3);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    q = fma(q, r, 1.0);
    return q;
  }
  __device__ __forceinline__ double exp_acc(double x, double acc) const {
    double s; double p = reduced(x, s);
    return fma(p, s, acc);
  }
  __device__ __forceinline__ double exp(double x) const {
    double s; double p = reduced(x, s);
    return p * s;
  }
};

}  // namespace dpgp

// ---------------------------------------------------------------------------------------------------
// 256-entry shared-memory table variant (EXPV = 4):  x = (256 e + j) ln2/256 + r, |r| <= ln2/512,
//   exp(x) = 2^e * T[j] * (1 + r + r^2/2 + r^3/6 + r^4/24),   T[j] = 2^(j/256) (correctly rounded, host-built).
// Truncation error r^5/120 <= 3.8e-17 relative; the table entry carries <= 1.1e-16; the single-word
// ln2/256 reduction adds |x / ln2| * 2.3e-17, the same as the polynomial variant (measured on a CPU
// emulation against expl: 2.2e-15 on [-60, 5]).
// FP64-pipe cost of  w * exp(x):  3 (reduction) + 3 (Horner) + 3 (T*w, T*r, final FMA) = 9 issues,
// vs 16 for the degree-11 polynomial; the table read is one LDS.64 (data-dependent address).
// Validity of the integer part is checked on the high word of the magic-number sum, so any x below
// about -1022 ln2 (including hugely negative sentinels) returns exactly 0.
namespace dpgp {

constexpr int kExpTabBits = 8;
constexpr int kExpTabSize = 1 << kExpTabBits;

struct ExpTabConst {
  static constexpr double MAGIC = 6755399441055744.0;            // 1.5 * 2^52
  static constexpr double L2E_S = 0x1.71547652b82fep+8;          // 256 / ln2
  static constexpr double NLN2_S = -0x1.62e42fefa39efp-9;        // -ln2 / 256
  static constexpr double C3 = 1.0 / 24.0, C2 = 1.0 / 6.0;
};

// Scaled table entry 2^e T[j] from the magic-number sum t = x * 256/ln2 + MAGIC (0 when out of range).
__device__ __forceinline__ double exp_tab_entry(const double* __restrict__ tab, double t) {
  const int lo = __double2loint(t), hi = __double2hiint(t);
  const int e = lo >> kExpTabBits;
  const double tj = tab[lo & (kExpTabSize - 1)];
  // t = MAGIC + k with |k| < 2^31  <=>  hi == 0x43380000 - (k < 0)
  const bool ok = (hi == 0x43380000 - (int)((unsigned)lo >> 31)) && (e >= -1022);
  const int thi = __double2hiint(tj) + (e << 20);
  return __hiloint2double(ok ? thi : 0, ok ? __double2loint(tj) : 0);
}

// out[k] = w[k] * exp(x[k]), K chains in lockstep.
template <int K>
__device__ __forceinline__ void exp_tab_scaled_k(const double* __restrict__ tab, const double (&x)[K], const double (&w)[K],
                                                 double (&out)[K]) {
  double t[K], r[K], q[K], T[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(x[k], ExpTabConst::L2E_S, ExpTabConst::MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) { T[k] = exp_tab_entry(tab, t[k]); t[k] -= ExpTabConst::MAGIC; }
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(t[k], ExpTabConst::NLN2_S, x[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(r[k], ExpTabConst::C3, ExpTabConst::C2);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], 0.5);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < K; ++k) T[k] *= w[k];
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = fma(T[k] * r[k], q[k], T[k]);
}

// acc[k] += exp(x[k]), K chains in lockstep (9 FP64 issues per chain).
template <int K>
__device__ __forceinline__ void exp_tab_acc_k(const double* __restrict__ tab, const double (&x)[K], double (&acc)[K]) {
  double t[K], r[K], q[K], T[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(x[k], ExpTabConst::L2E_S, ExpTabConst::MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) { T[k] = exp_tab_entry(tab, t[k]); t[k] -= ExpTabConst::MAGIC; }
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(t[k], ExpTabConst::NLN2_S, x[k]);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(r[k], ExpTabConst::C3, ExpTabConst::C2);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], 0.5);
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], 1.0);
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] += T[k];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = fma(T[k] * r[k], q[k], acc[k]);
}

}  // namespace dpgp
