// Everything of the objective that does not depend on N, in two single-CTA kernels (forward, backward):
//   * positive variables  softplus(raw)                                  (reference src/utils/types.py:52-57)
//   * phi = softmax(logits), rows repeated mask_size times               (src/models/dirichlet_process.py:39-51)
//   * the truncated stick-breaking DP objective -ELBO, six terms          (dirichlet_process.py:64-88; beta.py:8-19,
//     gamma.py:8-17, multinomial.py:8-16)
//   * the log-normal(0,1) hyper-prior on the atoms                        (dp_gp_lvm.py:96-98 / :603-605, log_normal.py:24-39)
//   * D-mode: gamma_d = phi gamma_atoms, alpha_d, beta_d                   (dp_gp_lvm.py:100-102)
// and the chain of the cotangents of the GP bound (d phi, d gamma, d alpha, d beta from bound_kernel / the statistics
// backward) back to the raw variables.  TensorFlow autodiff did this with a few hundred tiny ops per iteration; here it is
// the closed form (trigamma for the digamma terms).  At the reference's own problem sizes (N = 100..300) these ~300 launches
// were most of a training iteration.
//
// With A_t = psi(g1_t) - psi(g1_t + g2_t), B_t = psi(g2_t) - psi(g1_t + g2_t) (t < T-1), c = psi(w1) - log w2, rho = w1 / w2,
// Phi_t = sum_d phi_dt, Tail_t = sum_{j > t} Phi_j:
//   ELBO = sum_t (Phi_t A_t + Tail_t B_t) + (T-1) c + (rho - 1) sum_t B_t + s1 log s2 - lgamma(s1) + (s1 - 1) c - s2 rho
//          - sum phi log phi + sum_t H_Beta(g1_t, g2_t) + H_Gamma(w1, w2)
// All sums run in a fixed order (one thread per column / shared-memory tree): bitwise reproducible.
#pragma once
#include "common.cuh"
#include "special.cuh"

namespace dpgp {

constexpr int kSmallThreads = 512;
constexpr int kSmallMaxT = 256;

struct SmallParams {
  // raw variables
  const double* logits;          // [D / mask][T]
  const double* g1_raw; const double* g2_raw;      // [T-1]
  const double* w1_raw; const double* w2_raw;      // scalars
  const double* ga_raw;          // [T][Q]
  const double* aa_raw; const double* ba_raw;      // [T]
  // forward outputs
  double* phi;                   // [D][T]
  double* gamma; double* alpha; double* beta;      // [B][Q], [B], [B]   B = T (T-mode) or D (D-mode: phi-mixtures)
  double* scal;                  // [2]: DP objective (-ELBO), hyper-prior
  // backward inputs: cotangents of the GP bound gp (the objective is dp - gp - prior), grad_out of the objective
  const double* dphi;            // [D][T] or NULL (D-mode)
  const double* dgamma; const double* dalpha; const double* dbeta;   // [B][Q], [B], [B]
  const double* grad_out;        // [1] or NULL (= 1)
  // backward outputs (raw-variable gradients of the objective)
  double* dlogits; double* dg1_raw; double* dg2_raw; double* dw1_raw; double* dw2_raw;
  double* dga_raw; double* daa_raw; double* dba_raw;
  int d, t, q, mask, mode;       // mode 0 = T, 1 = D
  double s1, s2;
};

// shared layout (doubles): A[T] B[T] colsum[T] tailB[T] dPhi[T] red[64] | g1 g2 (T each) | atoms ga[T*Q] aa[T] ba[T]
struct SmallShared {
  double *A, *B, *col, *pre, *g1, *g2, *ga, *aa, *ba, *red, *tmp;
};
__device__ __forceinline__ SmallShared small_carve(double* sm, int t, int q) {
  SmallShared s;
  s.A = sm; s.B = s.A + t; s.col = s.B + t; s.pre = s.col + t; s.g1 = s.pre + t; s.g2 = s.g1 + t;
  s.aa = s.g2 + t; s.ba = s.aa + t; s.tmp = s.ba + t; s.red = s.tmp + t; s.ga = s.red + 64;
  return s;
}
__host__ __device__ inline size_t small_smem_bytes(int t, int q) { return ((size_t)10 * t + 64 + (size_t)t * q) * 8; }

// Fixed-order block sum, result broadcast to every thread.
__device__ __forceinline__ double block_sum_all(double v, double* red) {
  const double r = block_sum(v, red);
  __syncthreads();
  if (threadIdx.x == 0) red[32] = r;
  __syncthreads();
  return red[32];
}

// Common prologue of both kernels: positive variables, A / B, atoms into shared memory.
__device__ __forceinline__ void small_prologue(const SmallParams& p, const SmallShared& s) {
  const int tid = threadIdx.x, T = p.t;
  for (int t = tid; t < T; t += blockDim.x) {
    if (t < T - 1) {
      const double g1 = softplus_d(p.g1_raw[t]), g2 = softplus_d(p.g2_raw[t]);
      const double d12 = digamma_pos(g1 + g2);
      s.g1[t] = g1; s.g2[t] = g2; s.A[t] = digamma_pos(g1) - d12; s.B[t] = digamma_pos(g2) - d12;
    } else { s.g1[t] = 1.0; s.g2[t] = 1.0; s.A[t] = 0.0; s.B[t] = 0.0; }
    s.aa[t] = softplus_d(p.aa_raw[t]); s.ba[t] = softplus_d(p.ba_raw[t]);
  }
  for (int i = tid; i < T * p.q; i += blockDim.x) s.ga[i] = softplus_d(p.ga_raw[i]);
  __syncthreads();
}

// Row of phi: stable softmax of one logits row; returns through shared-free registers is impossible for runtime T, so the
// row is written to global phi and log phi is recomputed where needed from (logit - max - log sum).
static __global__ void __launch_bounds__(kSmallThreads) small_fwd_kernel(SmallParams p) {
  extern __shared__ __align__(16) double sm[];
  const SmallShared s = small_carve(sm, p.t, p.q);
  const int tid = threadIdx.x, NT = blockDim.x, T = p.t, D = p.d, depth = p.d / p.mask;
  small_prologue(p, s);
  // ---- phi rows and the q(Z) entropy  -sum phi log phi  over all D rows (repeated rows count, dp_gp_lvm.py:584-588).
  //      One warp per row, lanes over the T columns (the first version ran one THREAD per row and one thread per column
  //      sum: 67 us for D x T = 60 x 20, all of it serial latency).
  const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  double hz = 0.0;
  for (int r = warp; r < depth; r += nwarps) {
    const double* lg = p.logits + (size_t)r * T;
    double mx = -1.0e300;
    for (int t = lane; t < T; t += 32) mx = fmax(mx, lg[t]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double se = 0.0;
    for (int t = lane; t < T; t += 32) se += exp(lg[t] - mx);
    se = warp_sum(se);
    const double lse = log(se), inv = 1.0 / se;
    double h = 0.0;
    for (int t = lane; t < T; t += 32) {
      const double e = exp(lg[t] - mx), ph = e * inv;
      h -= ph * ((lg[t] - mx) - lse);
      for (int k = 0; k < p.mask; ++k) p.phi[((size_t)r * p.mask + k) * T + t] = ph;
    }
    h = warp_sum(h);
    if (lane == 0) hz += h * p.mask;
  }
  hz = block_sum_all(hz, s.red);
  __syncthreads();                                        // phi visible to the whole CTA (global writes + barrier)
  // ---- column sums Phi_t: one warp per column, lanes over the rows, fixed order
  for (int t = warp; t < T; t += nwarps) {
    double a = 0.0;
    for (int r = lane; r < depth; r += 32) a += p.phi[((size_t)r * p.mask) * T + t];
    a = warp_sum(a);
    if (lane == 0) s.col[t] = a * p.mask;
  }
  __syncthreads();
  // ---- the six ELBO terms: thread t < T - 1 evaluates its q(V) entropy and E[log p(Z|V)] summands (lgamma / digamma are the
  //      slow part: ~6 per t), fixed-order block sums combine them
  {
    if (tid == 0) { double tail = 0.0; for (int t = T - 2; t >= 0; --t) { tail += s.col[t + 1]; s.tmp[t] = tail; } }
    __syncthreads();
    double e1 = 0.0, sumB = 0.0, hv = 0.0;
    for (int t = tid; t < T - 1; t += NT) {
      e1 += s.col[t] * s.A[t] + s.tmp[t] * s.B[t];
      sumB += s.B[t];
      const double g1 = s.g1[t], g2 = s.g2[t], tot = g1 + g2;
      const double d12 = digamma_pos(tot), dg1 = s.A[t] + d12, dg2 = s.B[t] + d12;
      hv += lgamma(g1) + lgamma(g2) - lgamma(tot) - (g1 - 1.0) * dg1 - (g2 - 1.0) * dg2 + (tot - 2.0) * d12;
    }
    e1 = block_sum_all(e1, s.red); sumB = block_sum_all(sumB, s.red); hv = block_sum_all(hv, s.red);
    if (tid == 0) {
      const double w1 = softplus_d(*p.w1_raw), w2 = softplus_d(*p.w2_raw);
      const double psiw = digamma_pos(w1), c = psiw - log(w2), rho = w1 / w2;
      const double e2 = (T - 1.0) * c + (rho - 1.0) * sumB;
      const double e3 = p.s1 * log(p.s2) - lgamma(p.s1) + (p.s1 - 1.0) * c - p.s2 * rho;
      const double ha = w1 - log(w2) + lgamma(w1) + (1.0 - w1) * psiw;
      p.scal[0] = -(e1 + e2 + e3 + hz + hv + ha);
    }
    __syncthreads();
  }
  // ---- hyper-prior  sum log N(log x; 0, 1) - log x  over the atoms
  {
    double a = 0.0;
    const int na = T * p.q + 2 * T;
    for (int i = tid; i < na; i += NT) {
      const double x = i < T * p.q ? s.ga[i] : (i < T * p.q + T ? s.aa[i - T * p.q] : s.ba[i - T * p.q - T]);
      const double lx = log(x);
      a += -lx - 0.5 * (1.8378770664093453 + lx * lx);
    }
    a = block_sum(a, s.red);
    if (tid == 0) p.scal[1] = a;
  }
  // ---- kernel-batch hyper-parameters
  if (p.mode == 0) {
    for (int i = tid; i < T * p.q; i += NT) p.gamma[i] = s.ga[i];
    for (int t = tid; t < T; t += NT) { p.alpha[t] = s.aa[t]; p.beta[t] = s.ba[t]; }
  } else {
    for (int i = tid; i < D * (p.q + 2); i += NT) {
      const int d = i / (p.q + 2), j = i - d * (p.q + 2);
      const double* ph = p.phi + (size_t)d * T;
      double a = 0.0;
      if (j < p.q) { for (int t = 0; t < T; ++t) a = fma(ph[t], s.ga[t * p.q + j], a); p.gamma[(size_t)d * p.q + j] = a; }
      else if (j == p.q) { for (int t = 0; t < T; ++t) a = fma(ph[t], s.aa[t], a); p.alpha[d] = a; }
      else { for (int t = 0; t < T; ++t) a = fma(ph[t], s.ba[t], a); p.beta[d] = a; }
    }
  }
}

// Gradient of  objective = dp - gp - prior  w.r.t. the raw small variables, times grad_out.
static __global__ void __launch_bounds__(kSmallThreads) small_bwd_kernel(SmallParams p) {
  extern __shared__ __align__(16) double sm[];
  const SmallShared s = small_carve(sm, p.t, p.q);
  const int tid = threadIdx.x, NT = blockDim.x, T = p.t, D = p.d, Q = p.q, depth = p.d / p.mask;
  const double go = p.grad_out ? *p.grad_out : 1.0;
  small_prologue(p, s);
  // ---- column sums of phi (for the q(V) / q(alpha) terms) and prefix sums of B (d ELBO / d phi_dj needs sum_{t<j} B_t)
  const int lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
  for (int t = warp; t < T; t += nwarps) {                // one warp per column, lanes over the rows (as in the forward kernel)
    double a = 0.0;
    for (int r = lane; r < depth; r += 32) a += p.phi[((size_t)r * p.mask) * T + t];
    a = warp_sum(a);
    if (lane == 0) s.col[t] = a * p.mask;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0;
    for (int t = 0; t < T; ++t) { s.pre[t] = a; if (t < T - 1) a += s.B[t]; }       // pre[j] = sum_{t < j, t < T-1} B_t
  }
  __syncthreads();
  // ---- atoms: d objective / d atom = -(gp cotangent) - d prior, through softplus
  {
    const int na = T * Q + 2 * T;
    for (int i = tid; i < na; i += NT) {
      const int which = i < T * Q ? 0 : (i < T * Q + T ? 1 : 2);
      const int t = which == 0 ? i / Q : (which == 1 ? i - T * Q : i - T * Q - T), qq = which == 0 ? i - t * Q : 0;
      const double x = which == 0 ? s.ga[i] : (which == 1 ? s.aa[t] : s.ba[t]);
      double g = 0.0;                                       // d gp / d atom
      if (p.mode == 0) g = which == 0 ? p.dgamma[i] : (which == 1 ? p.dalpha[t] : p.dbeta[t]);
      else {
        for (int d = 0; d < D; ++d) {
          const double ph = p.phi[(size_t)d * T + t];
          g = fma(ph, which == 0 ? p.dgamma[(size_t)d * Q + qq] : (which == 1 ? p.dalpha[d] : p.dbeta[d]), g);
        }
      }
      const double lx = log(x), dprior = -(1.0 + lx) / x;
      const double raw = which == 0 ? p.ga_raw[i] : (which == 1 ? p.aa_raw[t] : p.ba_raw[t]);
      const double out = go * (-g - dprior) * sigmoid_d(raw);
      if (which == 0) p.dga_raw[i] = out; else if (which == 1) p.daa_raw[t] = out; else p.dba_raw[t] = out;
    }
  }
  // ---- q(V), q(alpha)
  if (tid < T - 1) {
    const int t = tid;                                    // T - 1 <= kSmallThreads is checked by the host
    double tail = 0.0;
    for (int j = t + 1; j < T; ++j) tail += s.col[j];
    const double g1 = s.g1[t], g2 = s.g2[t], tot = g1 + g2;
    const double t1 = trigamma_pos(g1), t2 = trigamma_pos(g2), t12 = trigamma_pos(tot);
    const double w1 = softplus_d(*p.w1_raw), w2 = softplus_d(*p.w2_raw), rho = w1 / w2;
    // d ELBO / d g1, g2
    const double de1 = s.col[t] * (t1 - t12) + (tail + rho - 1.0) * (-t12) + (-(g1 - 1.0) * t1 + (tot - 2.0) * t12);
    const double de2 = s.col[t] * (-t12) + (tail + rho - 1.0) * (t2 - t12) + (-(g2 - 1.0) * t2 + (tot - 2.0) * t12);
    p.dg1_raw[t] = go * (-de1) * sigmoid_d(p.g1_raw[t]);
    p.dg2_raw[t] = go * (-de2) * sigmoid_d(p.g2_raw[t]);
  }
  if (tid == 0) {
    const double w1 = softplus_d(*p.w1_raw), w2 = softplus_d(*p.w2_raw), rho = w1 / w2, tw = trigamma_pos(w1);
    double sumB = 0.0;
    for (int t = 0; t < T - 1; ++t) sumB += s.B[t];
    const double dw1 = (T - 1.0) * tw + sumB / w2 + (p.s1 - 1.0) * tw - p.s2 / w2 + 1.0 + (1.0 - w1) * tw;
    const double dw2 = -(T - 1.0) / w2 - rho / w2 * sumB - (p.s1 - 1.0) / w2 + p.s2 * rho / w2 - 1.0 / w2;
    *p.dw1_raw = go * (-dw1) * sigmoid_d(*p.w1_raw);
    *p.dw2_raw = go * (-dw2) * sigmoid_d(*p.w2_raw);
  }
  // ---- logits: per row of the (unrepeated) softmax, d objective / d phi summed over the mask_size repeated rows.
  //      One warp per row, lanes over the columns: dlogit_t = phi_t (a_t - sum_t' phi_t' a_t'),  a_t = d objective / d phi_t
  for (int r = warp; r < depth; r += nwarps) {
    const double* lg = p.logits + (size_t)r * T;
    double mx = -1.0e300;
    for (int t = lane; t < T; t += 32) mx = fmax(mx, lg[t]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    double se = 0.0;
    for (int t = lane; t < T; t += 32) se += exp(lg[t] - mx);
    se = warp_sum(se);
    const double lse = log(se), inv = 1.0 / se;
    double dot = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int t = lane; t < T; t += 32) {
        const double lph = (lg[t] - mx) - lse, ph = exp(lg[t] - mx) * inv;
        // d ELBO_dp / d phi (per repeated row) = A_t [t < T-1] + pre[t] - (log phi + 1)
        const double ddp = -((t < T - 1 ? s.A[t] : 0.0) + s.pre[t] - (lph + 1.0));       // d dp_objective / d phi
        double acc = 0.0;
        for (int k = 0; k < p.mask; ++k) {
          const int d = r * p.mask + k;
          double dgp;                                       // d gp / d phi_dt
          if (p.mode == 0) dgp = p.dphi[(size_t)d * T + t];
          else {
            dgp = p.dalpha[d] * s.aa[t] + p.dbeta[d] * s.ba[t];
            for (int qq = 0; qq < Q; ++qq) dgp = fma(p.dgamma[(size_t)d * Q + qq], s.ga[t * Q + qq], dgp);
          }
          acc += ddp - dgp;
        }
        if (pass == 0) dot = fma(ph, acc, dot);
        else p.dlogits[(size_t)r * T + t] = go * ph * (acc - dot);
      }
      if (pass == 0) dot = warp_sum(dot);
    }
  }
}

}  // namespace dpgp
