// Shared-memory Cholesky + triangular inverse of one M x M matrix per CTA, and a batched small-GEMM "job list" kernel:
// the building blocks of the M x M chain of the collapsed bound (bound.cuh; reference src/models/dp_gp_lvm.py:113-133 /
// :618-640: tf.cholesky + tf.matrix_triangular_solve per kernel-batch entry).
//
// Round 1 ran the whole chain as ONE CTA per kernel-batch entry over matrices in global scratch (~30 barrier-separated
// phases, 1.44 ms at M = 128: the named limiter of the 8-GPU run and the largest kernel of a small-shape training
// iteration).  Now the two inherently sequential pieces -- chol(K_uu + 1e-8 I) and chol(beta H + I), each followed by the
// inverse of the factor -- run with the matrix held in SHARED MEMORY (factor_kernel, M <= 144; a global-memory version
// covers M up to 256), and everything else is a dense product spread over the whole GPU by mm_jobs_kernel.
//
// factor_kernel, one CTA of 512 threads per matrix, blocked right-looking, NB = 16:
//   per block column:  warp 0 factors the 16 x 16 diagonal block in registers (lane <-> row, all indices static, shuffles
//                      broadcast the pivot row; one rsqrt per column);
//                      panel below:  L21 D^T = A21  by substitution, one thread per row, right-looking in registers;
//                      trailing update A22 -= L21 L21^T on the FP64 tensor cores (mma.sync m8n8k4, 16 warps).
//   inverse (in place): all diagonal blocks inverted at once (one warp each), then block columns right to left:
//                      X21 = -X22 (L21 D^-1)  -- two tensor-core products per block column.
// A non-positive pivot sets the flag of the matrix (first failing row + 1) and makes the factor NaN, so that the objective
// and every gradient of that evaluation turn NaN (tf.cholesky raises at this point; dpgp_check reports the location).
#pragma once
#include "common.cuh"

namespace dpgp {

constexpr int kFacNB = 16;
constexpr int kFacMaxM = 144;            // shared-memory path: (144 x 148 + 9 x 16 x 17 + 144 x 24) doubles = 212 KB
constexpr int kFacThreads = 512;
constexpr int kFacTLd = 24;              // leading dimension of the panel buffer (== 8 mod 16: conflict-free B fragments)

__host__ __device__ inline int fac_mq(int m) { return (m + kFacNB - 1) / kFacNB * kFacNB; }
__host__ __device__ inline int fac_ld(int m) { return fac_mq(m) + 4; }            // == 4 mod 16: conflict-free A fragments
__host__ __device__ inline size_t fac_smem_bytes(int m) {
  const int mq = fac_mq(m);
  return ((size_t)mq * fac_ld(m) + (size_t)(mq / kFacNB) * kFacNB * (kFacNB + 1) + (size_t)mq * kFacTLd) * sizeof(double);
}

__device__ __forceinline__ void dmma884_f(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// 16 x 16 diagonal block at d (leading dimension ld): Cholesky in place (lower part); the reciprocals of the diagonal of the
// factor go to invdiag[16].  One warp; lane i < 16 holds row i in registers, all indices static, shuffles broadcast the pivot
// row, one rsqrt per column.  Returns (to every lane) the index + 1 of the first non-positive pivot, 0 if none.
__device__ __forceinline__ int diag_chol(double* d, int ld, double* invdiag) {
  const int lane = threadIdx.x & 31, row = lane & 15;
  const unsigned full = 0xffffffffu;
  double r[kFacNB];
#pragma unroll
  for (int j = 0; j < kFacNB; ++j) r[j] = d[(size_t)row * ld + j];
  int bad = 0;
#pragma unroll
  for (int k = 0; k < kFacNB; ++k) {
    const double piv = __shfl_sync(full, r[k], k);
    if (!(piv > 0.0) && bad == 0) bad = k + 1;
    const double ri = (piv > 0.0) ? rsqrt(piv) : nan("");
    if (lane == k) invdiag[k] = ri;
    r[k] = (row == k) ? piv * ri : r[k] * ri;             // l_kk = sqrt(piv), l_ik = a_ik / l_kk
#pragma unroll
    for (int j = k + 1; j < kFacNB; ++j) {
      const double ljk = __shfl_sync(full, r[k], j);        // l_jk from the lane that owns row j
      r[j] = fma(-r[k], ljk, r[j]);                         // only entries with row >= j are used later
    }
  }
  if (lane < kFacNB) {
#pragma unroll
    for (int j = 0; j < kFacNB; ++j) if (j <= row) d[(size_t)row * ld + j] = r[j];
  }
  return bad;
}

// Inverse X = L^-1 of the 16 x 16 lower-triangular block at d into dinv [16][17].  One warp; lane i holds row i of L, and
// lane j <-> column j of X:  x_jj = 1 / l_jj,  x_ij = -(1 / l_ii) sum_{k=j}^{i-1} l_ik x_kj.
__device__ __forceinline__ void diag_inv(const double* d, int ld, double* dinv) {
  const int lane = threadIdx.x & 31, row = lane & 15;
  const unsigned full = 0xffffffffu;
  double r[kFacNB];
#pragma unroll
  for (int j = 0; j < kFacNB; ++j) r[j] = d[(size_t)row * ld + j];
  double myinv = 1.0;
#pragma unroll
  for (int j = 0; j < kFacNB; ++j) if (j == row) myinv = 1.0 / r[j];
  double x[kFacNB];
#pragma unroll
  for (int i = 0; i < kFacNB; ++i) {
    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
    for (int k = 0; k < i; ++k) {
      const double lik = __shfl_sync(full, r[k], i);        // l_ik (lane i holds row i)
      const double xk = (k >= row) ? x[k] : 0.0;
      if (k & 1) acc1 = fma(lik, xk, acc1); else acc0 = fma(lik, xk, acc0);
    }
    const double invi = __shfl_sync(full, myinv, i);
    x[i] = (i == row) ? invi : ((i > row) ? -invi * (acc0 + acc1) : 0.0);
  }
  if (lane < kFacNB) {
#pragma unroll
    for (int i = 0; i < kFacNB; ++i) dinv[i * (kFacNB + 1) + row] = x[i];
  }
}

// In-place blocked Cholesky followed by the in-place inverse of the factor; `a` is mq x mq (mq multiple of 16, leading
// dimension ld) in shared memory with the lower triangle valid; on return the lower triangle holds L^-1 (the strictly
// upper part is garbage).  logdet_l = sum log l_ii over the first m rows.  All kFacThreads threads call.
__device__ void chol_inverse_smem(double* a, int mq, int ld, int m, double* dinv, double* tbuf, int* bad_out, int bad_offset,
                                  double* logdet_l, double* red /* >= 33 doubles of shared memory */, double* lout /* [m][m] or NULL */) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3, nwarps = kFacThreads / 32;
  const int nblk = mq / kFacNB;
  __shared__ int s_bad;
  __shared__ double invd[kFacNB];
  if (tid == 0) s_bad = 0;
  __syncthreads();
  for (int jb = 0; jb < nblk; ++jb) {
    const int j0 = jb * kFacNB, r0 = j0 + kFacNB, nrows = mq - r0;
    if (warp == 0) {
      const int bad = diag_chol(a + (size_t)j0 * ld + j0, ld, invd);
      if (lane == 0 && bad && s_bad == 0) s_bad = j0 + bad;
    }
    __syncthreads();
    if (nrows > 0) {
      // ---- panel: L21 D^T = A21 by substitution, one thread per row, right-looking in registers (in place: a thread reads
      //      and writes only its own row)
      if (tid < nrows) {
        double* ai = a + (size_t)(r0 + tid) * ld + j0;
        const double* dblk = a + (size_t)j0 * ld + j0;
        double v[kFacNB];
#pragma unroll
        for (int c = 0; c < kFacNB; ++c) v[c] = ai[c];
#pragma unroll
        for (int c = 0; c < kFacNB; ++c) {
          const double x = v[c] * invd[c];
          v[c] = x;
#pragma unroll
          for (int c2 = c + 1; c2 < kFacNB; ++c2) v[c2] = fma(-x, dblk[(size_t)c2 * ld + c], v[c2]);
        }
#pragma unroll
        for (int c = 0; c < kFacNB; ++c) ai[c] = v[c];
      }
      __syncthreads();
      // ---- trailing update (lower triangle by 8 x 8 tiles): A22 -= L21 L21^T
      const int nt = nrows / 8, ntiles = nt * (nt + 1) / 2;
      for (int t = warp; t < ntiles; t += nwarps) {
        int ti = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);      // float estimate, corrected exactly below
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        const int tj = t - ti * (ti + 1) / 2;
        const double* pa = a + (size_t)(r0 + 8 * ti + lr) * ld + j0 + lc;
        const double* pb = a + (size_t)(r0 + 8 * tj + lr) * ld + j0 + lc;
        double c[2] = {0.0, 0.0};
#pragma unroll
        for (int ks = 0; ks < kFacNB / 4; ++ks) dmma884_f(c, pa[4 * ks], pb[4 * ks]);
        double* pc = a + (size_t)(r0 + 8 * ti + lr) * ld + r0 + 8 * tj + 2 * lc;
        pc[0] -= c[0]; pc[1] -= c[1];
      }
      __syncthreads();
    }
  }
  // log-determinant of the factor (the diagonal still holds l_ii here); fixed-order sum
  {
    double part = 0.0;
    for (int i = tid; i < m; i += kFacThreads) part += log(a[(size_t)i * ld + i]);
    part = block_sum(part, red);
    if (tid == 0) *logdet_l = part;
    if (lout)
      for (int idx = tid; idx < m * m; idx += kFacThreads) { const int i = idx / m, j = idx - i * m; lout[idx] = (j <= i) ? a[(size_t)i * ld + j] : 0.0; }
  }
  __syncthreads();
  // ---- inverse in place.  First all diagonal blocks at once (independent: one warp each; mq / 16 <= 9 <= 16 warps) ...
  if (warp < nblk) diag_inv(a + (size_t)(warp * kFacNB) * ld + warp * kFacNB, ld, dinv + (size_t)warp * kFacNB * (kFacNB + 1));
  __syncthreads();
  // ... then block columns from right to left:  X21 = -X22 (L21 Dj^-1),  Xjj = Dj^-1
  for (int jb = nblk - 1; jb >= 0; --jb) {
    const int j0 = jb * kFacNB, r0 = j0 + kFacNB, nrows = mq - r0;
    const double* dj = dinv + (size_t)jb * kFacNB * (kFacNB + 1);
    if (nrows > 0) {
      // T[i][c] = sum_{k >= c} L21[i][k] X[k][c]
      for (int idx = tid; idx < nrows * kFacNB; idx += kFacThreads) {
        const int i = idx >> 4, c = idx & 15;
        const double* ai = a + (size_t)(r0 + i) * ld + j0;
        double s = 0.0;
        for (int k = c; k < kFacNB; ++k) s = fma(ai[k], dj[k * (kFacNB + 1) + c], s);
        tbuf[(size_t)i * kFacTLd + c] = s;
      }
    }
    __syncthreads();
    // diagonal block of the inverse (the panel below was consumed into tbuf)
    // (zeros above its diagonal: later block columns multiply whole 8 x 8 tiles of X22)
    if (tid < kFacNB * kFacNB) {
      const int i = tid >> 4, c = tid & 15;
      a[(size_t)(j0 + i) * ld + j0 + c] = (c <= i) ? dj[i * (kFacNB + 1) + c] : 0.0;
    }
    if (nrows > 0) {
      // X21 (nrows x 16) = -X22 (lower triangular, already inverted) * T : output tiles of 8 rows x 8 columns
      const int nt = nrows / 8;
      for (int t = warp; t < nt * 2; t += nwarps) {
        const int ti = t >> 1, tc = t & 1;
        double c[2] = {0.0, 0.0};
        const double* pa = a + (size_t)(r0 + 8 * ti + lr) * ld + r0 + lc;
        const double* pb = tbuf + (size_t)lc * kFacTLd + 8 * tc + lr;
        const int kend = 8 * ti + 8;                       // X22[i][k] = 0 for k > i
        for (int k0 = 0; k0 < kend; k0 += 4) dmma884_f(c, pa[k0], pb[(size_t)k0 * kFacTLd]);
        double* pc = a + (size_t)(r0 + 8 * ti + lr) * ld + j0 + 8 * tc + 2 * lc;
        pc[0] = -c[0]; pc[1] = -c[1];
      }
    }
    __syncthreads();
  }
  if (tid == 0 && s_bad && *bad_out == 0) *bad_out = s_bad + bad_offset;
}

// ------------------------------------------------------------------------------------------------------------------
// factor_kernel: builds the matrix of kernel-batch entry b in shared memory, factors and inverts it, writes the inverse
// of the factor (lower triangular, zero above the diagonal) to out [B][M][M], and a few scalars.
//   mode 0:  K_uu + 1e-8 I from (z, gamma, alpha)                    (rbf_kernel.py:58-93; same expansion as the reference)
//   mode 1:  A = beta_b H_b + I  from hmat [B][M][M]                 (dp_gp_lvm.py:630-633); also tr H.
//            H comes out of two triangular solves and is symmetric only up to rounding (~kappa eps |H|).  The factorisation reads
//            its lower triangle; that triangle is mirrored into the upper one here, because the closed-form cotangents use
//            H as a full matrix and must differentiate the function that was actually evaluated: with the two triangles left
//            different, the z gradient at kappa = 1e9 was off by 5.8e-6 instead of 1e-7 (profiles/r02_bound_cotangent_forms.txt).
// scal [B][4] (mode 1): { sum log diag(L_A), tr H, sum of squares of L_A^-1 (= tr A^-1), 0 }
struct FactorParams {
  const double* z; const double* gamma; const double* alpha;   // mode 0
  double* hmat; const double* beta;                            // mode 1 (the upper triangle of hmat is overwritten by the mirror of the lower)
  double* out;            // [B][M][M]  inverse of the factor
  double* lout;           // [B][M][M]  the factor itself (lower triangular, zeros above), or NULL
  double* scal;           // [B][4] or NULL
  int* bad;               // [B]
  int mode, m, q, bad_offset;
};

// sg[k] = sqrt(gamma_k), precomputed once per matrix (an FP64 sqrt per term was most of the matrix build)
__device__ __forceinline__ double kuu_entry(const double* z, const double* sg, double alpha, int i, int j, int q) {
  double xi = 0, xj = 0, xx = 0;
  for (int k = 0; k < q; ++k) {
    const double u = sg[k] * z[i * q + k], v = sg[k] * z[j * q + k];
    xi = fma(u, u, xi); xj = fma(v, v, xj); xx = fma(u, v, xx);
  }
  double kv = alpha * exp(-0.5 * xi - 0.5 * xj + xx);
  if (i == j) kv += kJitter;
  return kv;
}

__global__ void __launch_bounds__(kFacThreads, 1) factor_kernel(FactorParams p) {
  extern __shared__ __align__(16) double fsm[];
  __shared__ double red[40];
  const int b = blockIdx.x, tid = threadIdx.x, M = p.m, mq = fac_mq(M), ld = fac_ld(M);
  double* a = fsm;
  double* dinv = a + (size_t)mq * ld;
  double* tbuf = dinv + (size_t)(mq / kFacNB) * kFacNB * (kFacNB + 1);
  const size_t mm = (size_t)M * M;
  double trh_part = 0.0;
  if (p.mode == 0) {
    __shared__ double sg[kMaxQ];
    if (tid < p.q) sg[tid] = sqrt(p.gamma[(size_t)b * p.q + tid]);
    __syncthreads();
    const double* gam = sg;
    const double alpha = p.alpha[b];
    for (int idx = tid; idx < mq * mq; idx += kFacThreads) {
      const int i = idx / mq, j = idx - i * mq;
      double v = (i == j) ? 1.0 : 0.0;                     // identity padding keeps the padded factor trivial
      if (i < M && j < M) v = (j <= i) ? kuu_entry(p.z, gam, alpha, i, j, p.q) : 0.0;
      a[(size_t)i * ld + j] = v;
    }
  } else {
    double* hm = p.hmat + (size_t)b * mm;
    const double beta = p.beta[b];
    for (int idx = tid; idx < mq * mq; idx += kFacThreads) {
      const int i = idx / mq, j = idx - i * mq;
      double v = (i == j) ? 1.0 : 0.0;
      if (i < M && j < M) {
        const double hv = hm[(size_t)(i >= j ? i : j) * M + (i >= j ? j : i)];      // lower triangle
        if (i < j) hm[(size_t)i * M + j] = hv;                                       // mirrored into the upper one
        v = beta * hv + (i == j ? 1.0 : 0.0);
        if (i == j) trh_part += hv;
      }
      a[(size_t)i * ld + j] = v;
    }
  }
  __syncthreads();
  __shared__ double s_logdet;
  chol_inverse_smem(a, mq, ld, M, dinv, tbuf, &p.bad[b], p.bad_offset, &s_logdet, red, p.lout ? p.lout + (size_t)b * mm : nullptr);
  __syncthreads();
  double* out = p.out + (size_t)b * mm;
  double sq_part = 0.0;
  for (int idx = tid; idx < M * M; idx += kFacThreads) {
    const int i = idx / M, j = idx - i * M;
    const double v = (j <= i) ? a[(size_t)i * ld + j] : 0.0;
    out[idx] = v;
    sq_part = fma(v, v, sq_part);
  }
  if (p.scal) {
    const double trh = block_sum(trh_part, red);
    const double sq = block_sum(sq_part, red);
    if (tid == 0) { double* s = p.scal + (size_t)b * 4; s[0] = s_logdet; s[1] = trh; s[2] = sq; s[3] = 0.0; }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// X = L^-1 B by blocked forward substitution, parallel over kernel-batch entries and over blocks of 32 right-hand-side
// columns (the columns of a triangular solve are independent).  This is how the reference forms H = L^-1 Psi2 L^-T and
// C = L_A^-1 L^-1 P (tf.matrix_triangular_solve, dp_gp_lvm.py:118-121,132-133 / :621-624,638-639); products with the
// explicit inverse agree with it only to ~100 kappa eps, substitution to ~kappa eps (measured, profiles/r02_bound.md).
// A CTA (4 warps) keeps its M x 32 slab of X in shared memory; per block of 16 rows: the 16 x k row block of L is
// prefetched with cp.async (double-buffered), the update X_I -= L_I,<I X_<I runs on the FP64 tensor cores, and warp 0
// substitutes through the 16 x 16 diagonal block, one lane per column.
//   B(i,j) = src[b*ss + i*si + j*sj]   (any transposition),   X -> out[b*so + i*ldo + j]
constexpr int kTsCols = 32, kTsNB = 16, kTsLdX = 40, kTsMaxJobs = 2;
struct TrsmJob { const double* l; const double* src; double* out; long long sl, ss, so; int si, sj, ldo, nc, blk0; };
struct TrsmJobs { TrsmJob j[kTsMaxJobs]; int count, m; };
__host__ __device__ inline int ts_ldl(int m) { return fac_mq(m) + 4; }
__host__ __device__ inline size_t ts_smem_bytes(int m) {
  return ((size_t)fac_mq(m) * kTsLdX + 2 * (size_t)kTsNB * ts_ldl(m)) * sizeof(double);
}
__global__ void __launch_bounds__(128) trsm_cols_kernel(TrsmJobs jobs) {
  extern __shared__ __align__(16) double tsm[];
  int ji = 0;
  while (ji + 1 < jobs.count && (int)blockIdx.x >= jobs.j[ji + 1].blk0) ++ji;
  const TrsmJob& J = jobs.j[ji];
  const int M = jobs.m, mq = fac_mq(M), ldl = ts_ldl(M), nblk = mq / kTsNB;
  const int b = blockIdx.y, c0 = (blockIdx.x - J.blk0) * kTsCols;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  double* Xs = tsm;                                    // [mq][kTsLdX]
  double* Ls = Xs + (size_t)mq * kTsLdX;               // [2][16][ldl]
  const double* L = J.l + (size_t)b * J.sl;
  const double* src = J.src + (size_t)b * J.ss;
  // right-hand side slab; the walk follows the unit stride of the source
  for (int idx = tid; idx < mq * kTsCols; idx += 128) {
    int i, j;
    if (J.sj == 1) { i = idx >> 5; j = idx & 31; } else { j = idx / mq; i = idx - j * mq; }
    Xs[(size_t)i * kTsLdX + j] = (i < M && c0 + j < J.nc) ? src[(size_t)i * J.si + (size_t)(c0 + j) * J.sj] : 0.0;
  }
  auto prefetch = [&](int ib) {                        // rows 16 ib .. 16 ib + 15, columns 0 .. 16 ib + 15 of L
    double* dst = Ls + (size_t)(ib & 1) * kTsNB * ldl;
    const int i0 = ib * kTsNB, w = i0 + kTsNB;
    for (int idx = tid; idx < kTsNB * w; idx += 128) {
      const int r = idx / w, k = idx - r * w, i = i0 + r;
      if (i < M && k < M) cp_async8(dst + (size_t)r * ldl + k, L + (size_t)i * M + k);
      else dst[(size_t)r * ldl + k] = (i == k) ? 1.0 : 0.0;                 // identity padding
    }
    cp_async_commit();
  };
  prefetch(0);
  for (int ib = 0; ib < nblk; ++ib) {
    const int i0 = ib * kTsNB;
    cp_async_wait<0>();
    __syncthreads();                                   // L row block ib (and, first time, the slab) visible to everybody
    const double* Lr = Ls + (size_t)(ib & 1) * kTsNB * ldl;
    if (ib + 1 < nblk) prefetch(ib + 1);               // lands while this block row is processed
    if (i0 > 0) {
      // X_I (16 x 32) -= L[I, 0:i0] X[0:i0, :]; warp <-> 8 columns, both 8-row tiles
      double c[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
      const double* pa0 = Lr + (size_t)lr * ldl + lc;
      const double* pa1 = pa0 + 8 * ldl;
      const double* pb = Xs + (size_t)lc * kTsLdX + 8 * warp + lr;
#pragma unroll 4
      for (int k0 = 0; k0 < i0; k0 += 4) {
        const double bf = pb[(size_t)k0 * kTsLdX];
        dmma884_f(c[0], pa0[k0], bf);
        dmma884_f(c[1], pa1[k0], bf);
      }
      double* px = Xs + (size_t)(i0 + lr) * kTsLdX + 8 * warp + 2 * lc;
      px[0] -= c[0][0]; px[1] -= c[0][1];
      px[8 * kTsLdX] -= c[1][0]; px[8 * kTsLdX + 1] -= c[1][1];
    }
    __syncthreads();
    if (warp == 0) {
      // substitution through the diagonal block, lane <-> column, right-looking: as soon as x_i is known every later row is
      // updated (independent FMAs), so the dependent chain is one multiply + one FMA per row instead of a dot product
      double sv[kTsNB];
#pragma unroll
      for (int i = 0; i < kTsNB; ++i) sv[i] = Xs[(size_t)(i0 + i) * kTsLdX + lane];
      const double myinv = 1.0 / Lr[(size_t)(lane & 15) * ldl + i0 + (lane & 15)];
#pragma unroll
      for (int i = 0; i < kTsNB; ++i) {
        const double x = sv[i] * __shfl_sync(0xffffffffu, myinv, i);
        sv[i] = x;
#pragma unroll
        for (int r = i + 1; r < kTsNB; ++r) sv[r] = fma(-Lr[(size_t)r * ldl + i0 + i], x, sv[r]);
      }
#pragma unroll
      for (int i = 0; i < kTsNB; ++i) Xs[(size_t)(i0 + i) * kTsLdX + lane] = sv[i];
    }
    // (the barrier at the top of the next iteration orders these writes before the next update)
  }
  __syncthreads();
  double* out = J.out + (size_t)b * J.so;
  for (int idx = tid; idx < M * kTsCols; idx += 128) {
    const int i = idx >> 5, j = idx & 31;
    if (c0 + j < J.nc) out[(size_t)i * J.ldo + c0 + j] = Xs[(size_t)i * kTsLdX + j];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Batched small dense products on the FP64 tensor cores.  One launch evaluates up to kMmMaxJobs independent products for
// every kernel-batch entry:  C_b (n x m) = alpha * op(A_b) diag(w_b) op(B_b), operands addressed by element strides so that
// every transposition is free:  A(i,k) = a[b*sa + i*ai + k*ak],  B(k,j) = bm[b*sb + k*bk + j*bj],  C(i,j) = c[b*sc + i*ldc + j],
// w_k = w[b*sw + k*wk] (optional).
// A CTA (4 warps) owns a 32 x 32 tile of C; K is walked in slabs of 32 staged in shared memory (k contiguous for both
// operands: conflict-free fragment loads at leading dimension 36).  Triangular structure only trims the k range of a tile
// (the zeros of the operands are stored).  sym: only tiles on or below the diagonal are computed and mirrored.
constexpr int kMmMaxJobs = 4;
constexpr int kMmTile = 32, kMmKc = 32, kMmLd = 36;
enum { MM_A_LOWER = 1 /* A(i,k) = 0 for k > i */, MM_A_UPPER = 2 /* k < i */, MM_B_KLEJ = 4 /* B(k,j) = 0 for k > j */,
       MM_B_KGEJ = 8 /* B(k,j) = 0 for k < j */, MM_SYM = 16 /* C symmetric: lower tiles computed, mirrored */ };
struct MmJob {
  const double* a; const double* bm; double* c; const double* w;
  long long sa, sb, sc, sw;
  int ai, ak, bk, bj, ldc, wk;     // wk: element stride of the weights over k
  int n, m, k;
  int flags;
  double alpha;
  int tiles_n, tiles_m, tile0;
};
struct MmJobs { MmJob j[kMmMaxJobs]; int count; };

__global__ void __launch_bounds__(128) mm_jobs_kernel(MmJobs jobs) {
  __shared__ __align__(16) double As[kMmTile * kMmLd];
  __shared__ __align__(16) double Bs[kMmTile * kMmLd];
  int ji = 0;
  while (ji + 1 < jobs.count && (int)blockIdx.x >= jobs.j[ji + 1].tile0) ++ji;
  const MmJob& J = jobs.j[ji];
  const int t = blockIdx.x - J.tile0, ti = t / J.tiles_m, tj = t - ti * J.tiles_m;
  if ((J.flags & MM_SYM) && tj > ti) return;
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lr = lane >> 2, lc = lane & 3;
  const int i0 = ti * kMmTile, j0 = tj * kMmTile;
  const double* A = J.a + (size_t)b * J.sa;
  const double* B = J.bm + (size_t)b * J.sb;
  const double* W = J.w ? J.w + (size_t)b * J.sw : nullptr;
  int klo = 0, khi = J.k;
  if (J.flags & MM_A_LOWER) khi = min(khi, i0 + kMmTile);
  if (J.flags & MM_A_UPPER) klo = max(klo, i0);
  if (J.flags & MM_B_KLEJ) khi = min(khi, j0 + kMmTile);
  if (J.flags & MM_B_KGEJ) klo = max(klo, j0);
  klo = klo / kMmKc * kMmKc;
  const int wr = (warp >> 1) * 16, wc = (warp & 1) * 16;
  double acc[2][2][2];
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 2; ++y) { acc[x][y][0] = 0.0; acc[x][y][1] = 0.0; }
  for (int k0 = klo; k0 < khi; k0 += kMmKc) {
    // stage A (32 x 32) and B (32 x 32) slabs, k contiguous in shared memory; the global walk follows the unit stride
#pragma unroll
    for (int e = 0; e < (kMmTile * kMmKc) / 128; ++e) {
      const int idx = tid + e * 128;
      int i, k;
      if (J.ak == 1) { i = idx >> 5; k = idx & 31; } else { k = idx >> 5; i = idx & 31; }
      double v = 0.0;
      if (i0 + i < J.n && k0 + k < J.k) {
        v = A[(size_t)(i0 + i) * J.ai + (size_t)(k0 + k) * J.ak];
        if (W) v *= W[(size_t)(k0 + k) * J.wk];
      }
      As[i * kMmLd + k] = v;
      int j, k2;
      if (J.bk == 1) { j = idx >> 5; k2 = idx & 31; } else { k2 = idx >> 5; j = idx & 31; }
      double u = 0.0;
      if (j0 + j < J.m && k0 + k2 < J.k) u = B[(size_t)(k0 + k2) * J.bk + (size_t)(j0 + j) * J.bj];
      Bs[j * kMmLd + k2] = u;
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < kMmKc / 4; ++ks) {
      double af[2], bf[2];
#pragma unroll
      for (int x = 0; x < 2; ++x) af[x] = As[(wr + 8 * x + lr) * kMmLd + 4 * ks + lc];
#pragma unroll
      for (int y = 0; y < 2; ++y) bf[y] = Bs[(wc + 8 * y + lr) * kMmLd + 4 * ks + lc];
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int y = 0; y < 2; ++y) dmma884_f(acc[x][y], af[x], bf[y]);
    }
    __syncthreads();
  }
  double* C = J.c + (size_t)b * J.sc;
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = i0 + wr + 8 * x + lr, j = j0 + wc + 8 * y + 2 * lc + e;
        if (i < J.n && j < J.m) {
          const double v = J.alpha * acc[x][y][e];
          C[(size_t)i * J.ldc + j] = v;
          if ((J.flags & MM_SYM) && ti != tj) C[(size_t)j * J.ldc + i] = v;
        }
      }
}

}  // namespace dpgp
