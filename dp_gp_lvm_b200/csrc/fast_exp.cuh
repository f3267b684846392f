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
// Shared-memory table variants (EXPV = 4, 5, 6):  x = (S e + j) ln2/S + r, |r| <= ln2/(2S),
//   exp(x) = 2^e * T[j] * (1 + r q(r)),  q = 1 + r/2 + ... + r^(d-1)/d!,  T[j] = 2^(j/S) (correctly rounded, host-built)
//   EXPV 4: S = 256, d = 4 (truncation r^5/120 <= 3.8e-17)    9 FP64 issues for acc += exp(x); table read ~6 wavefronts
//   EXPV 5: S =  64, d = 5 (r^6/720  <= 3.5e-17)             10 issues; entries share 4-way -> ~3.5 wavefronts
//   EXPV 6: S =  32, d = 6 (r^7/5040 <= 3.4e-18)             11 issues; at most 2-way bank conflicts
// (the degree-11 polynomial costs 15 issues and no table read).  The table entry carries <= 1.1e-16 relative
// error and the single-word ln2/S reduction adds |x / ln2| * 2.3e-17, as in the polynomial variant (CPU
// emulation against expl: 2.2e-15 on [-60, 5]).  The table read is one LDS.64 at a data-dependent address.
// Validity of the integer part is checked on the high word of the magic-number sum, so any x below about
// -1022 ln2 (including hugely negative sentinels) returns exactly 0.
namespace dpgp {

constexpr int kExpTabSize = 256;                 // shared-memory doubles reserved for the table (largest variant)
__host__ __device__ constexpr int exp_tab_bits(int expv) { return expv == 8 ? 11 : expv == 6 ? 5 : expv == 5 ? 6 : 8; }
// EXPV 8 (forward kernel only, where shared memory has room): 2 048 entries, |r| <= ln2 / 4096, degree-3 polynomial
// (truncation r^4 / 24 <= 3.4e-17): 8 FP64 issues per accumulated exp instead of 9.
constexpr int kExpTabSizeFwd = 2048;
// (Round 2 experiment, removed: the 32-entry table replicated 16 times -- entry j of copy c at [16 j + c], a lane reads copy
// lane & 15 -- makes the data-dependent read conflict-free (2 wavefronts instead of ~5.5) but needs the degree-6 polynomial:
// forward 38.9 -> 42.3 ms, fused backward 95.1 -> 98.4 ms at 262 144 rows.  Both kernels pay more for two FP64 issues per exp
// than they gain from 3.5 shared-memory wavefronts.)

template <int BITS>
struct ExpTabConst {
  static constexpr double MAGIC = 6755399441055744.0;                           // 1.5 * 2^52
  static constexpr double L2E_S = 0x1.71547652b82fep+0 * (double)(1 << BITS);   // S / ln2
  static constexpr double NLN2_S = -0x1.62e42fefa39efp-1 / (double)(1 << BITS); // -ln2 / S
};

// Scaled table entry 2^e T[j] from the magic-number sum t = x * S/ln2 + MAGIC (0 when out of range).
// DPGP_EXP_CLAMPED (default): t = 1.5 2^52 + k holds the 64-bit integer k = (hi(t) - 0x43380000) : lo(t), so one funnel shift
// gives k >> BITS for every |k| < 2^(31 + BITS) (|x| < 1.4e9, the documented domain) and one IMNMX clamps it at -1023:
// 6 integer instructions per exp instead of 11 (two ISETP, two SEL and the sign fix-up of the validated form go away).
// Results below 2^-1022 come out as 2^-1023 T[j] (a denormal <= 1.2e-308) instead of exactly 0.
// -DDPGP_EXP_CLAMPED=0 restores the validated form.
#ifndef DPGP_EXP_CLAMPED
#define DPGP_EXP_CLAMPED 1
#endif
template <int BITS>
__device__ __forceinline__ double exp_tab_entry(const double* __restrict__ tab, double t) {
#if DPGP_EXP_CLAMPED
  {
    const int lo_ = __double2loint(t), hk_ = __double2hiint(t) - 0x43380000;
    const int e_ = max((int)__funnelshift_r((unsigned)lo_, (unsigned)hk_, BITS), -1023);
    const double tj_ = tab[lo_ & ((1 << BITS) - 1)];
    return __hiloint2double(__double2hiint(tj_) + (e_ << 20), __double2loint(tj_));
  }
#endif
  const int lo = __double2loint(t), hi = __double2hiint(t);
  const int e = lo >> BITS;
  const double tj = tab[lo & ((1 << BITS) - 1)];
  // t = MAGIC + k with |k| < 2^31  <=>  hi == 0x43380000 - (k < 0)
  const bool ok = (hi == 0x43380000 - (int)((unsigned)lo >> 31)) && (e >= -1022);
  const int thi = __double2hiint(tj) + (e << 20);
  return __hiloint2double(ok ? thi : 0, ok ? __double2loint(tj) : 0);
}

// q(r) = 1 + r/2 + r^2/6 + ... (degree BITS-dependent), K chains in lockstep
template <int BITS, int K>
__device__ __forceinline__ void exp_tab_poly(const double (&r)[K], double (&q)[K]) {
  constexpr int DEG = BITS >= 11 ? 3 : BITS >= 8 ? 4 : BITS == 6 ? 5 : 6;
  constexpr double c[7] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0};
#pragma unroll
  for (int k = 0; k < K; ++k) q[k] = fma(r[k], c[DEG], c[DEG - 1]);
#pragma unroll
  for (int j = DEG - 2; j >= 1; --j) {
#pragma unroll
    for (int k = 0; k < K; ++k) q[k] = fma(q[k], r[k], c[j]);
  }
}

// out[k] = w[k] * exp(x[k]), K chains in lockstep.
template <int BITS, int K>
__device__ __forceinline__ void exp_tab_scaled_k(const double* __restrict__ tab, const double (&x)[K], const double (&w)[K],
                                                 double (&out)[K]) {
  using C = ExpTabConst<BITS>;
  double t[K], r[K], q[K], T[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(x[k], C::L2E_S, C::MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) { T[k] = exp_tab_entry<BITS>(tab, t[k]); t[k] -= C::MAGIC; }
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(t[k], C::NLN2_S, x[k]);
  exp_tab_poly<BITS, K>(r, q);
#pragma unroll
  for (int k = 0; k < K; ++k) T[k] *= w[k];
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = fma(T[k] * r[k], q[k], T[k]);
}

// acc[k] += exp(x[k]), K chains in lockstep.
template <int BITS, int K>
__device__ __forceinline__ void exp_tab_acc_k(const double* __restrict__ tab, const double (&x)[K], double (&acc)[K]) {
  using C = ExpTabConst<BITS>;
  double t[K], r[K], q[K], T[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = fma(x[k], C::L2E_S, C::MAGIC);
#pragma unroll
  for (int k = 0; k < K; ++k) { T[k] = exp_tab_entry<BITS>(tab, t[k]); t[k] -= C::MAGIC; }
#pragma unroll
  for (int k = 0; k < K; ++k) r[k] = fma(t[k], C::NLN2_S, x[k]);
  exp_tab_poly<BITS, K>(r, q);
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] += T[k];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = fma(T[k] * r[k], q[k], acc[k]);
}

}  // namespace dpgp
