// The M x M chain of the collapsed bound for one kernel-batch entry b per CTA, forward + closed-form
// backward (reference src/models/dp_gp_lvm.py:113-145 (D-mode) / :618-667 (T-mode); backward replaces
// tf.gradients).  Operation order of the forward follows the reference ("stable" form,
// test/unittests/bgplvm_unittests.py:85-112):
//     L = chol(K_uu + 1e-8 I);  H = L^-1 Psi2 L^-T (two triangular solves);  A = beta H + I;  L_A = chol(A)
//     C = L_A^-1 L^-1 P;   q_c = ||C[:,c]||^2
// Closed forms, Sigma = K + beta Psi2, S = Sigma^-1 = R^T R with R = L_A^-1 L^-1:
//     F_b = n_b [ 1/2 (N log beta + beta (tr H - alpha N)) - logdet L_A ] - 1/2 beta sum_c w_c yy_c + 1/2 beta^2 sum_c w_c q_c
//     dF/dPsi2 = 1/2 n beta (K^-1 - S) - 1/2 beta^3 W          W = U diag(w) U^T,  U = S P
//     dF/dK    = n (-1/2 beta K^-1 Psi2 K^-1 - 1/2 S + 1/2 K^-1) - 1/2 beta^2 W
//   evaluated WITHOUT the subtractions: with Linv = L^-1, Lainv = L_A^-1, R = Lainv Linv, G = Lainv H Linv and
//   A^-1 - I = -beta H A^-1,    K^-1 - S = beta R^T G,    K^-1 - S - beta K^-1 Psi2 K^-1 = -beta^2 G^T G,
//   so dF/dPsi2 = 1/2 n beta^2 R^T G - 1/2 beta^3 W and dF/dK = -1/2 n beta^2 G^T G - 1/2 beta^2 W.  (K^-1 and S are nearly
//   equal when beta H << I, e.g. few rows per inducing point, and the dense matrices K^-1, S, K^-1 Psi2 K^-1 are then not
//   needed at all; dpgp_bound_factors forms K^-1 and S on demand for the prediction paths.)  H must be exactly symmetric
//   for these forms to be the derivative of what was evaluated: factor_kernel mirrors the triangle it factors.
//   Measured (profiles/r02_bound_cotangent_forms.txt): z gradient at kappa = 1.2e9 within 1.0e-7 (round 1: 5.2e-7).
//     dF/dP    = beta^2 U diag(w)
//     dF/dbeta = n [ N/(2 beta) + 1/2 (tr H - alpha N) - 1/2 tr(S Psi2) ] + beta sum w q - 1/2 beta^2 tr(W Psi2) - 1/2 sum w yy
//     dF/dalpha (direct) = -1/2 n beta N
//     dF/dw_c  = 1/2 (N log beta + beta (tr H - alpha N)) - logdet L_A - 1/2 beta yy_c + 1/2 beta^2 q_c
// The chain of dF/dK into Z, gamma, alpha (and of dD from psi2) is zchain_kernel below.
#pragma once
#include "common.cuh"
#include "factor.cuh"

namespace dpgp {

// Evaluation order (dpgp_api.cu: dpgp_kuu_factor / dpgp_bound), all of it parallel over the kernel batch:
//   [side stream, needs only Z / gamma / alpha]   Linv = (chol(K_uu + 1e-8 I))^-1   (factor_kernel, shared memory)
//                                                 Kinv = Linv^T Linv                  (mm_jobs_kernel)
//   [after the all-reduce of the statistics]      T1 = L^-1 Psi2, C1 = L^-1 P;  H = (L^-1 T1^T)^T    (trsm_cols_kernel x 2: the
//                                                 reference's own triangular solves, parallel over columns)
//                                                 L_A = chol(beta H + I), Lainv = L_A^-1, log det, tr H, tr A^-1   (factor_kernel)
//                                                 Cm = L_A^-1 C1 (trsm_cols_kernel);  R = Lainv Linv;  T2 = Linv^T H
//                                                 S = R^T R;  KPK = T2 Linv
//                                                 U = R^T Cm
//                                                 PU = Psi2 U;  W = U diag(w) U^T
//                                                 bound_out_kernel (closed forms above), bound_finish_kernel, zchain_kernel
// H and C are formed by triangular solves with L and L_A as in the reference (the forward value rests on them); the inverse
// factors serve the cotangents, which need Kinv and S as dense matrices anyway.  Every step is spread over B x (M/32) or
// B x (M/32)^2 CTAs; the only single-CTA pieces are the two factorisations, which run out of shared memory.
// Round 1 ran the whole chain as one CTA per kernel-batch entry in global scratch: 1.44 ms at M = 128, B = 10.
//
// ---- dense helpers of the global-memory fallback (M > kFacMaxM): all threads of the CTA cooperate -----------------

__device__ __forceinline__ void dmma884_b(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// out (n x m, ldo)  =  sign * op(A) op(B) [+ out if accumulate],  inner dimension kk, optional weights colw[k] on A.
// op(A)[i][k] = ta ? a[k*lda+i] : a[i*lda+k];  op(B)[k][j] = tb ? b[j*ldb+k] : b[k*ldb+j].
// A warp owns blocks of 16 x 32 outputs (2 x 4 DMMA tiles); everything outside the matrices reads as zero.
__device__ void gemm_dmma(const double* a, int lda, bool ta, const double* b, int ldb, bool tb, double* out, int ldo,
                          int n, int m, int kk, const double* colw, double sign = 1.0, bool accumulate = false) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lr = lane >> 2, lc = lane & 3;
  const int bn = (n + 15) / 16, bm = (m + 31) / 32;
  for (int blk = warp; blk < bn * bm; blk += nwarps) {
    const int i0 = (blk / bm) * 16, j0 = (blk % bm) * 32;
    double c[2][4][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int t = 0; t < 4; ++t) { c[r][t][0] = 0.0; c[r][t][1] = 0.0; }
    for (int k0 = 0; k0 < kk; k0 += 4) {
      const int k = k0 + lc;
      const bool kin = k < kk;
      const double wk = (colw && kin) ? colw[k] : 1.0;
      double af[2], bf[4];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = i0 + r * 8 + lr;
        af[r] = (kin && i < n) ? (ta ? a[(size_t)k * lda + i] : a[(size_t)i * lda + k]) * wk : 0.0;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = j0 + t * 8 + lr;
        bf[t] = (kin && j < m) ? (tb ? b[(size_t)j * ldb + k] : b[(size_t)k * ldb + j]) : 0.0;
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int t = 0; t < 4; ++t) dmma884_b(c[r][t], af[r], bf[t]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = i0 + r * 8 + lr, j = j0 + t * 8 + 2 * lc + e;
          if (i < n && j < m) {
            double* o = out + (size_t)i * ldo + j;
            *o = accumulate ? fma(sign, c[r][t][e], *o) : sign * c[r][t][e];
          }
        }
  }
  __syncthreads();
}

// In-place lower Cholesky of the symmetric matrix a (n x n, ld), left-looking by columns; the dot product of a row is
// split over 8 lanes.  Returns via *bad the first non-positive pivot index + 1 (0 = ok).  Upper triangle is zeroed.
__device__ void chol_lower(double* a, int n, int ld, int* bad, double* colbuf) {
  const int tid = threadIdx.x, T = blockDim.x, sub = tid & 7, grp = tid >> 3, ngrp = T >> 3;
  for (int j = 0; j < n; ++j) {
    // s_i = a[i][j] - sum_{k<j} L[i][k] L[j][k]   for i >= j
    const double* lj = a + (size_t)j * ld;
    for (int base = j; base < n; base += ngrp) {            // uniform trip count: every lane reaches the shuffles
      const int i = base + grp;
      const bool live = i < n;
      const double* li = a + (size_t)(live ? i : j) * ld;
      double s = 0.0;
      for (int k = sub; k < j; k += 8) s = fma(li[k], lj[k], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (live && sub == 0) colbuf[i] = li[j] - s;
    }
    __syncthreads();
    const double piv = colbuf[j];
    if (!(piv > 0.0)) { if (tid == 0 && *bad == 0) *bad = j + 1; }
    const double d = piv > 0.0 ? sqrt(piv) : nan("");      // NaN factor -> NaN objective and gradients (tf.cholesky raises here)
    for (int i = j + tid; i < n; i += T) a[(size_t)i * ld + j] = (i == j) ? d : colbuf[i] / d;
    __syncthreads();
  }
  for (int idx = tid; idx < n * n; idx += T) { int i = idx / n, k = idx % n; if (k > i) a[(size_t)i * ld + k] = 0.0; }
  __syncthreads();
}

// X = L^-1 B (forward substitution), B n x nc.  If bt: B is given transposed (nc x n, ldb).  Blocked by 16 rows: the
// update B_I - L_IJ X_J of a block row is a tensor-core contraction, the 16 x 16 diagonal solve is one thread per column.
constexpr int kTrsmNB = 16;
__device__ void trsm_lower_impl(const double* l, int n, int ldl, const double* b, int nc, int ldb, bool bt, double* x, int ldx) {
  const int tid = threadIdx.x, T = blockDim.x;
  for (int idx = tid; idx < n * nc; idx += T) {
    const int i = idx / nc, c = idx % nc;
    x[(size_t)i * ldx + c] = bt ? b[(size_t)c * ldb + i] : b[(size_t)i * ldb + c];
  }
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += kTrsmNB) {
    const int nb = min(kTrsmNB, n - i0);
    // X_I -= L[I, 0:i0] X[0:i0, :]
    if (i0 > 0) gemm_dmma(l + (size_t)i0 * ldl, ldl, false, x, ldx, false, x + (size_t)i0 * ldx, ldx, nb, nc, i0, nullptr, -1.0, true);
    for (int c = tid; c < nc; c += T) {
      for (int i = 0; i < nb; ++i) {
        const double* li = l + (size_t)(i0 + i) * ldl + i0;
        double s = x[(size_t)(i0 + i) * ldx + c];
        for (int k = 0; k < i; ++k) s = fma(-li[k], x[(size_t)(i0 + k) * ldx + c], s);
        x[(size_t)(i0 + i) * ldx + c] = s / li[i];
      }
    }
    __syncthreads();
  }
}
__device__ void trsm_lower(const double* l, int n, int ldl, const double* b, int nc, int ldb, double* x, int ldx) {
  trsm_lower_impl(l, n, ldl, b, nc, ldb, false, x, ldx);
}
// same with B given transposed: solves L X = B^T where bt is nc x n (ldbt)
__device__ void trsm_lower_bt(const double* l, int n, int ldl, const double* bt, int nc, int ldbt, double* x, int ldx) {
  trsm_lower_impl(l, n, ldl, bt, nc, ldbt, true, x, ldx);
}

// out = op(A) op(B): ta/tb select transposes; out n x m, inner dimension kk.  Optional column weights.
__device__ void gemm_small(const double* a, int lda, bool ta, const double* b, int ldb, bool tb,
                           double* out, int ldo, int n, int m, int kk, const double* colw /* weights on k */) {
  gemm_dmma(a, lda, ta, b, ldb, tb, out, ldo, n, m, kk, colw);
}

// ---- fallback for M > kFacMaxM (the matrix does not fit in shared memory): the same factor + inverse in global scratch -------
// work [B][M][M] receives the factor, tmp [B][M][M] the identity; out = L^-1 by blocked forward substitution.
struct FactorGlobalParams {
  FactorParams f;      // f.hmat is written (mirror), hence not const-qualified below
  double* work; double* tmp;
};
__global__ void __launch_bounds__(512) factor_global_kernel(FactorGlobalParams g) {
  __shared__ double red[40];
  __shared__ double colbuf[kMaxM];
  const FactorParams& p = g.f;
  const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x, M = p.m;
  const size_t mm = (size_t)M * M;
  double* a = g.work + (size_t)b * mm;
  double* eye = g.tmp + (size_t)b * mm;
  double* out = p.out + (size_t)b * mm;
  double trh_part = 0.0;
  __shared__ double sg[kMaxQ];
  if (p.mode == 0 && tid < p.q) sg[tid] = sqrt(p.gamma[(size_t)b * p.q + tid]);
  __syncthreads();
  for (int idx = tid; idx < M * M; idx += T) {
    const int i = idx / M, j = idx - i * M;
    double v;
    if (p.mode == 0) v = kuu_entry(p.z, sg, p.alpha[b], i, j, p.q);
    else {
      // lower triangle of H; its mirror is written back (see factor_kernel)
      const double hv = p.hmat[(size_t)b * mm + (i >= j ? idx : (size_t)j * M + i)];
      if (i < j) p.hmat[(size_t)b * mm + idx] = hv;
      v = p.beta[b] * hv + (i == j ? 1.0 : 0.0); if (i == j) trh_part += hv;
    }
    a[idx] = v;
    eye[idx] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  int before = p.bad[b];
  chol_lower(a, M, M, &p.bad[b], colbuf);
  if (tid == 0 && before == 0 && p.bad[b] != 0) p.bad[b] += p.bad_offset;
  double ld_part = 0.0;
  for (int i = tid; i < M; i += T) ld_part += log(a[(size_t)i * M + i]);
  if (p.lout) for (int idx = tid; idx < M * M; idx += T) p.lout[(size_t)b * mm + idx] = a[idx];      // chol_lower zeroed the upper part
  trsm_lower(a, M, M, eye, M, M, out, M);
  double sq_part = 0.0;
  for (int idx = tid; idx < M * M; idx += T) { const double v = out[idx]; sq_part = fma(v, v, sq_part); }
  if (p.scal) {
    const double ldet = block_sum(ld_part, red);
    const double trh = block_sum(trh_part, red);
    const double sq = block_sum(sq_part, red);
    if (tid == 0) { double* s = p.scal + (size_t)b * 4; s[0] = ldet; s[1] = trh; s[2] = sq; s[3] = 0.0; }
  }
}

// ---- last stage of the chain: scalars and cotangents of one kernel-batch entry from the dense factors ----------------
// With Linv = L^-1, Lainv = L_A^-1, R = Lainv Linv (so S = R^T R = (K + beta Psi2)^-1, Kinv = Linv^T Linv):
//   Cm = L_A^-1 L^-1 P,  q_c = ||Cm[:,c]||^2,  U = R^T Cm = S P,  PU = Psi2 U,  W = U diag(w) U^T,  G2 = Lainv H,  G = G2 Linv,
//   R^T G and G^T G
// have been formed by mm_jobs_kernel; this kernel evaluates the closed forms at the top of this file.
//   blockIdx.x == 0 : F_b, dF/dbeta, dF/dalpha (direct), dF/dw_c (T-mode)
//   blockIdx.x >= 1 : dF/dPsi2, dF/dK (M x M) and dF/dP (M x C), grid-stride over the entries
struct BoundOutParams {
  const double* rtg; const double* gtg; const double* wmat; const double* g2; const double* lainv;   // [B][M][M]
  const double* cm; const double* u; const double* pu;                             // [B][M][C]
  const double* scal;                                                              // [B][4] from factor_kernel (mode 1)
  const double* yy; const double* alpha; const double* beta; const double* wgt;
  double* fb; double* dpsi2; double* dp; double* dk; double* dbeta; double* dalpha_direct; double* dwgt;
  int64_t n_total; int d, m, b, mode, ncols;
};
__global__ void __launch_bounds__(256) bound_out_kernel(BoundOutParams p) {
  __shared__ double red[40];
  __shared__ double sc[4];
  const int b = blockIdx.y, tid = threadIdx.x, T = blockDim.x, M = p.m, C = p.ncols;
  const size_t mm = (size_t)M * M, mc = (size_t)M * C;
  const double alpha = p.alpha[b], beta = p.beta[b], N = (double)p.n_total;
  const int col0 = (p.mode == 1) ? b : 0;
  // n_b = sum_c w_c (fixed-order block sum; every block of this b computes the same value)
  double nb_part = 0.0;
  for (int c = tid; c < C; c += T) nb_part += p.wgt ? p.wgt[(size_t)(col0 + c) * p.b + b] : 1.0;
  double nb = block_sum(nb_part, red);
  if (tid == 0) sc[0] = nb;
  __syncthreads();
  nb = sc[0];
  if (blockIdx.x > 0) {
    const double b2 = beta * beta, b3 = b2 * beta;
    const double* rtg = p.rtg + (size_t)b * mm; const double* gtg = p.gtg + (size_t)b * mm; const double* wm = p.wmat + (size_t)b * mm;
    double* dpsi2 = p.dpsi2 + (size_t)b * mm; double* dk = p.dk + (size_t)b * mm;
    const size_t stride = (size_t)(gridDim.x - 1) * T;
    for (size_t idx = (size_t)(blockIdx.x - 1) * T + tid; idx < mm; idx += stride) {
      const double w = wm[idx];
      dpsi2[idx] = 0.5 * nb * b2 * rtg[idx] - 0.5 * b3 * w;       // 1/2 n beta (Kinv - S) - 1/2 beta^3 W
      dk[idx] = -0.5 * nb * b2 * gtg[idx] - 0.5 * b2 * w;         // n (-1/2 beta KPK - 1/2 S + 1/2 Kinv) - 1/2 beta^2 W
    }
    const double* u = p.u + (size_t)b * mc;
    double* dp = p.dp + (size_t)b * mc;
    for (size_t idx = (size_t)(blockIdx.x - 1) * T + tid; idx < mc; idx += stride) {
      const int c = (int)(idx % C);
      const double w = p.wgt ? p.wgt[(size_t)(col0 + c) * p.b + b] : 1.0;
      dp[idx] = b2 * u[idx] * w;
    }
    return;
  }
  const double* sca = p.scal + (size_t)b * 4;
  const double logdet = sca[0], trh = sca[1];
  const double* cm = p.cm + (size_t)b * mc; const double* u = p.u + (size_t)b * mc; const double* pu = p.pu + (size_t)b * mc;
  const double base_w = 0.5 * (N * log(beta) + beta * (trh - alpha * N)) - logdet;
  double sq = 0, syy = 0, swp = 0;      // sum w q, sum w yy, sum w u^T Psi2 u
  {
    // columns in chunks of 32 (lanes over the columns: coalesced), the rows of a chunk split over the 8 warps, partials
    // combined in fixed order by warp 0
    __shared__ double cpart[8][32][2];
    const int lane = tid & 31, warp = tid >> 5, nw = T >> 5;
    for (int cb = 0; cb < C; cb += 32) {
      const int c = cb + lane;
      double qc = 0, up = 0;
      if (c < C)
        for (int i = warp; i < M; i += nw) { const double v = cm[(size_t)i * C + c]; qc = fma(v, v, qc); up = fma(u[(size_t)i * C + c], pu[(size_t)i * C + c], up); }
      cpart[warp][lane][0] = qc; cpart[warp][lane][1] = up;
      __syncthreads();
      if (warp == 0 && c < C) {
        qc = 0; up = 0;
        for (int w = 0; w < nw; ++w) { qc += cpart[w][lane][0]; up += cpart[w][lane][1]; }
        const double w = p.wgt ? p.wgt[(size_t)(col0 + c) * p.b + b] : 1.0, yyc = p.yy[col0 + c];
        sq = fma(w, qc, sq); syy = fma(w, yyc, syy); swp = fma(w, up, swp);
        if (p.dwgt && p.mode == 0) p.dwgt[(size_t)c * p.b + b] = base_w - 0.5 * beta * yyc + 0.5 * beta * beta * qc;
      }
      __syncthreads();
    }
  }
  // tr(S Psi2) = tr(A^-1 H) = sum_ij (Lainv H)_ij Lainv_ij   (no cancellation; (M - tr A^-1) / beta loses digits for beta H << I)
  double trsp = 0.0;
  {
    const double* g2 = p.g2 + (size_t)b * mm; const double* la = p.lainv + (size_t)b * mm;
    double t4[4] = {0.0, 0.0, 0.0, 0.0};
    size_t idx = tid;
    for (; idx + 3 * (size_t)T < mm; idx += 4 * (size_t)T) {
#pragma unroll
      for (int e = 0; e < 4; ++e) t4[e] = fma(g2[idx + e * (size_t)T], la[idx + e * (size_t)T], t4[e]);
    }
    for (; idx < mm; idx += T) t4[0] = fma(g2[idx], la[idx], t4[0]);
    trsp = (t4[0] + t4[1]) + (t4[2] + t4[3]);
  }
  trsp = block_sum(trsp, red); if (tid == 0) sc[3] = trsp;
  sq = block_sum(sq, red); if (tid == 0) sc[1] = sq;
  syy = block_sum(syy, red); if (tid == 0) sc[2] = syy;
  swp = block_sum(swp, red);
  if (tid == 0) {
    sq = sc[1]; syy = sc[2]; trsp = sc[3];
    p.fb[b] = nb * base_w - 0.5 * beta * syy + 0.5 * beta * beta * sq;
    p.dbeta[b] = nb * (N / (2.0 * beta) + 0.5 * (trh - alpha * N) - 0.5 * trsp) + beta * sq - 0.5 * beta * beta * swp - 0.5 * syy;
    p.dalpha_direct[b] = -0.5 * nb * beta * N;
  }
}

// gp = -1/2 N D log(2 pi) + sum_b F_b - 1/2 (kl0 + kl1 - N Q); also fills the cotangents of yy and kl.
struct BoundFinishParams {
  const double* fb; const double* kl; const double* beta; const double* wgt;
  double* gp; double* dyy; double* dkl;
  int64_t n_total; int d, q, b, mode;
};
__global__ void bound_finish_kernel(BoundFinishParams p) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0;
    for (int b = 0; b < p.b; ++b) s += p.fb[b];
    const double N = (double)p.n_total;
    *p.gp = -0.5 * N * p.d * 1.8378770664093453 + s - 0.5 * (p.kl[0] + p.kl[1] - N * p.q);
    p.dkl[0] = -0.5; p.dkl[1] = -0.5;
  }
  for (int d = threadIdx.x; d < p.d; d += blockDim.x) {
    double a = 0;
    if (p.mode == 0) { for (int b = 0; b < p.b; ++b) a += p.wgt[(size_t)d * p.b + b] * p.beta[b]; }
    else a = p.beta[d];
    p.dyy[d] = -0.5 * a;
  }
}

// Chain of dF/dK (through K_uu) and of dD (psi2 pair side) into Z, gamma, alpha.
//   K0 = alpha exp(-1/2 sum_q g_q d_q^2), d = z_m - z_m'
//   dalpha += sum dK K0 / alpha ;  dgamma_q += -1/2 sum dK K0 d_q^2 ;  dz_mq += sum_m' (dK+dK^T)_mm' (-g_q d_q) K0
//   dz_mq += sum_{m'} 2 d_q ddsym[m,m',q]
// Grid (ceil(M / 8), B): a CTA of 8 warps owns 8 rows m of one kernel-batch entry, one warp per row, lanes over the columns
// c; K0 and its exponent are evaluated once per (m, c).  dz rows are complete per warp (fixed-order warp sums); the gamma /
// alpha sums leave the CTA as one partial per row block, summed in fixed order by bound_fin_kernel.  (Round 1: one CTA per
// kernel-batch entry, 82 us at M = 128 -- on the critical path of every evaluation and replicated on every rank.)
constexpr int kZcRows = 8;
struct ZChainParams {
  const double* dk;      // [B,M,M] or NULL
  const double* ddsym;   // [B,M,M,QP] or NULL
  const double* z; const double* gamma; const double* alpha;
  double* dz_b;          // [B,M,Q]  per-b contribution
  double* part;          // [B][row blocks][Q + 1]: partial sums of dgamma (Q) and of sum dK K0 (1), or NULL
  int q, qp, m, b;
};
// QMAX: register bound on Q (16 or 32); shared memory: M * Q doubles (dynamic).
template <int QMAX>
__global__ void __launch_bounds__(kZcRows * 32) zchain_kernel(ZChainParams p) {
  __shared__ double wpart[kZcRows][QMAX + 1];
  extern __shared__ __align__(16) double zs[];
  const int b = blockIdx.y, rb = blockIdx.x, tid = threadIdx.x, T = blockDim.x, M = p.m, Q = p.q;
  const int lane = tid & 31, warp = tid >> 5;
  const double alpha = p.alpha[b];
  for (int i = tid; i < M * Q; i += T) zs[i] = p.z[i];
  double gam[QMAX];
#pragma unroll
  for (int q = 0; q < QMAX; ++q) gam[q] = (q < Q) ? p.gamma[b * Q + q] : 0.0;
  __syncthreads();
  double da = 0;
  double dg[QMAX];
#pragma unroll
  for (int q = 0; q < QMAX; ++q) dg[q] = 0;
  const double* dk = p.dk ? p.dk + (size_t)b * M * M : nullptr;
  const double* dd = p.ddsym ? p.ddsym + (size_t)b * M * M * p.qp : nullptr;
  const int m = rb * kZcRows + warp;
  if (m < M) {
    double acc[QMAX];
#pragma unroll
    for (int q = 0; q < QMAX; ++q) acc[q] = 0;
    for (int c = lane; c < M; c += 32) {
      double dq[QMAX];
      double e = 0;
#pragma unroll
      for (int q = 0; q < QMAX; ++q) {
        dq[q] = (q < Q) ? zs[m * Q + q] - zs[c * Q + q] : 0.0;
        e = fma(gam[q] * dq[q], dq[q], e);
      }
      if (dk) {
        const double k0 = alpha * exp(-0.5 * e);
        const double gk = dk[(size_t)m * M + c] * k0;
        da += gk;
        const double gs = (c == m) ? 0.0 : (dk[(size_t)m * M + c] + dk[(size_t)c * M + m]) * k0;
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
          dg[q] = fma(-0.5 * gk, dq[q] * dq[q], dg[q]);
          acc[q] = fma(-gam[q] * dq[q], gs, acc[q]);
        }
      }
      if (dd && c != m) {
        const double* row = dd + ((size_t)m * M + c) * p.qp;
#pragma unroll
        for (int q = 0; q < QMAX; ++q)
          if (q < Q) acc[q] = fma(2.0 * dq[q], row[q], acc[q]);
      }
    }
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
      const double v = warp_sum(acc[q]);
      if (lane == 0 && q < Q) p.dz_b[((size_t)b * M + m) * Q + q] = v;
    }
  }
  if (!p.part) return;
#pragma unroll
  for (int q = 0; q < QMAX; ++q) { const double v = warp_sum(dg[q]); if (lane == 0) wpart[warp][q] = v; }
  da = warp_sum(da);
  if (lane == 0) wpart[warp][QMAX] = da;
  __syncthreads();
  if (tid <= Q) {
    const int q = tid < Q ? tid : QMAX;
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kZcRows; ++w) v += wpart[w][q];
    p.part[((size_t)b * gridDim.x + rb) * (Q + 1) + tid] = v;
  }
}

}  // namespace dpgp
