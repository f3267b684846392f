// psi1 statistic (reference src/kernels/rbf_kernel.py:135-161) contracted with Y on the fly, the
// column sums sum_n y^2 / KL sums, and the psi1 half of the backward.
//
//   log psi1_nm = log alpha - 1/2 sum_q [ w1_nq (mu_nq - z_mq)^2 + log(g s_nq + 1) ],  w1 = g/(g s + 1)
//   P[b] = Psi1[b]^T Y(:, cols(b))      T-mode: all D columns;  D-mode: the single column b
// Psi1 [B,N,M] is never written to HBM by the bound path (10 GB at the headline shape): each CTA builds
// a [rows x M] tile in shared memory and contracts it immediately.
#pragma once
#include "common.cuh"

namespace dpgp {

constexpr int kP1Rows = 32;      // rows of q(X) per tile
constexpr int kP1Cols = 64;      // columns of Y per contraction tile

// per-(n,q) quantities of one tile; lc[n] = log alpha - 1/2 sum_q log(g s + 1)
template <int QP>
__device__ __forceinline__ void psi1_row_terms(const double* mu, const double* s, const double* gamma, double alpha,
                                               int64_t n0, int nc, int Q, int b,
                                               double (*w1)[QP], double (*mus)[QP], double (*ld)[QP], double* lc) {
  for (int idx = threadIdx.x; idx < kP1Rows * QP; idx += blockDim.x) {
    int n = idx / QP, q = idx % QP;
    double w = 0, m_ = 0, l = 0;
    if (n < nc && q < Q) {
      double g = gamma[b * Q + q], sv = s[(n0 + n) * Q + q];
      double den = fma(g, sv, 1.0);
      w = g / den; l = log(den); m_ = mu[(n0 + n) * Q + q];
    }
    w1[n][q] = w; mus[n][q] = m_; ld[n][q] = l;
  }
  __syncthreads();
  if (threadIdx.x < kP1Rows) {
    double a = 0;
#pragma unroll
    for (int q = 0; q < QP; ++q) a += ld[threadIdx.x][q];
    lc[threadIdx.x] = log(alpha) - 0.5 * a;
  }
  __syncthreads();
}

// psi1 tile [kP1Rows][mp] into shared memory (rows >= nc and columns >= M are zero).
// thread <-> (m, row phase): blockDim/mp threads share a column and take interleaved rows, two rows in flight.
template <int QP>
__device__ __forceinline__ void psi1_tile(const double* z, int M, int mp, int Q, int nc,
                                          const double (*w1)[QP], const double (*mus)[QP], const double* lc,
                                          double* tile) {
  const int nparts = max(1, (int)blockDim.x / mp);
  for (int idx = threadIdx.x; idx < mp * nparts; idx += blockDim.x) {
    const int m = idx % mp, part = idx / mp;
    double zm[QP];
#pragma unroll
    for (int q = 0; q < QP; ++q) zm[q] = (m < M && q < Q) ? z[m * Q + q] : 0.0;
    for (int n = part; n < kP1Rows; n += 2 * nparts) {
      const int n2 = n + nparts;
      double a0 = 0, a1 = 0;
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        const double d0 = mus[n][q] - zm[q];
        a0 = fma(w1[n][q] * d0, d0, a0);
        if (n2 < kP1Rows) { const double d1 = mus[n2][q] - zm[q]; a1 = fma(w1[n2][q] * d1, d1, a1); }
      }
      tile[n * mp + m] = (n < nc && m < M) ? exp_fast(fmax(fma(-0.5, a0, lc[n]), -1.0e8)) : 0.0;
      if (n2 < kP1Rows) tile[n2 * mp + m] = (n2 < nc && m < M) ? exp_fast(fmax(fma(-0.5, a1, lc[n2]), -1.0e8)) : 0.0;
    }
  }
}

struct Psi1FwdParams {
  const double* mu; const double* s; const double* y; const double* z; const double* gamma; const double* alpha;
  double* part;          // [grid*2][mp*cpad]   per-CTA partial of P for (at most two) clusters
  int* tags;
  double* psi1_out;      // optional [B,N,M] materialisation (API surface / tests); NULL on the bound path
  int64_t n; int d, q, m, mp, b, mode, ncols, cpad, nseg; int64_t nchunks;
};

// T = 256 threads.  Register tile: 4 inducing points x 4 columns of Y per thread.  When all columns fit one
// column tile (ncols <= 64) and there are at most 2 tiles per thread (mp <= 128) the P accumulators stay in
// registers across all chunks of a cluster (PERSIST); otherwise each chunk adds its tile into the CTA's
// private partial in global memory.
template <int QP, bool PERSIST>
__global__ void __launch_bounds__(256, 2) psi1_fwd_kernel(Psi1FwdParams p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double w1[kP1Rows][QP], mus[kP1Rows][QP], ld[kP1Rows][QP], lc[kP1Rows];
  double* tile = sm;                                  // [kP1Rows][mp]
  double* yt = tile + kP1Rows * p.mp;                 // [kP1Cols][kP1Rows]  (transposed: column-major tile)
  const int tid = threadIdx.x, T = blockDim.x;
  const int64_t items = p.nchunks * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  if (lo >= hi) return;
  const int mtiles = p.mp / 4;
  const int nct = (p.ncols + kP1Cols - 1) / kP1Cols;
  const int ntiles = mtiles * (kP1Cols / 4);
  int cur_b = -1, seg = 0;
  double* mypart = nullptr;
  double pacc[2][4][4];
#pragma unroll
  for (int k = 0; k < 2; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) pacc[k][i][j] = 0.0;

  auto flush = [&]() {
    if (PERSIST) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int t = tid + k * T;
        if (t < ntiles) {
          const int m0 = (t % mtiles) * 4, c0 = (t / mtiles) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (c0 + j < p.ncols) mypart[(size_t)(m0 + i) * p.cpad + c0 + j] = pacc[k][i][j];
              pacc[k][i][j] = 0.0;
            }
        }
      }
    }
    if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = cur_b;
    ++seg;
  };

  for (int64_t item = lo; item < hi; ++item) {
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * kP1Rows;
    const int nc = (int)min((int64_t)kP1Rows, p.n - n0);
    if (b != cur_b) {
      if (cur_b >= 0) flush();
      cur_b = b;
      mypart = p.part + ((size_t)blockIdx.x * p.nseg + seg) * p.mp * p.cpad;
      if (!PERSIST) for (int i = tid; i < p.mp * p.cpad; i += T) mypart[i] = 0.0;
    }
    __syncthreads();
    psi1_row_terms<QP>(p.mu, p.s, p.gamma, p.alpha[b], n0, nc, p.q, b, w1, mus, ld, lc);
    psi1_tile<QP>(p.z, p.m, p.mp, p.q, nc, w1, mus, lc, tile);
    __syncthreads();
    if (p.psi1_out) {
      for (int i = tid; i < nc * p.m; i += T) {
        int n = i / p.m, m = i % p.m;
        p.psi1_out[((int64_t)b * p.n + n0 + n) * p.m + m] = tile[n * p.mp + m];
      }
    }
    const int col0 = (p.mode == 1) ? b : 0;
    for (int ct = 0; ct < nct; ++ct) {
      const int cbase = ct * kP1Cols;
      const int cw = min(kP1Cols, p.ncols - cbase);
      if (ct > 0) __syncthreads();
      for (int i = tid; i < kP1Cols * kP1Rows; i += T) {
        int n = i / kP1Cols, c = i % kP1Cols;          // coalesced over c
        double v = 0.0;
        if (n < nc && c < cw) v = p.y[(n0 + n) * p.d + col0 + cbase + c];
        yt[c * kP1Rows + n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        for (int t = tid + k * T; t < ntiles; t += 2 * T) {
          const int m0 = (t % mtiles) * 4, c0 = (t / mtiles) * 4;
          if (c0 >= cw) continue;
          double acc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = PERSIST ? pacc[k][i][j] : 0.0;
#pragma unroll 4
          for (int n = 0; n < kP1Rows; ++n) {
            const double2 a01 = *reinterpret_cast<const double2*>(tile + n * p.mp + m0);
            const double2 a23 = *reinterpret_cast<const double2*>(tile + n * p.mp + m0 + 2);
            const double av[4] = {a01.x, a01.y, a23.x, a23.y};
            double yv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) yv[j] = yt[(c0 + j) * kP1Rows + n];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], yv[j], acc[i][j]);
          }
          if (PERSIST) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) pacc[k][i][j] = acc[i][j];
          } else {
            // this CTA owns `mypart`; the (m,c) tile is owned by exactly one thread: plain read-modify-write
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (cbase + c0 + j < p.ncols) mypart[(size_t)(m0 + i) * p.cpad + cbase + c0 + j] += acc[i][j];
          }
        }
      }
    }
  }
  flush();
}

// Tensor-core version of the contraction P[b] += Psi1_tile^T Y_tile (north_star item 2: "the phi-weighted psi1^T Y
// contraction on FP64 DMMA tensor cores").  Same tiling, partial slots and tags as psi1_fwd_kernel<QP, true>; the
// 4x4 register tiles (6 shared-memory loads per 16 FMAs) become m8n8k4 DMMA tiles whose fragments are read from
// conflict-free layouts (leading dimensions = 4 mod 16 doubles): 10 loads per 16 DMMAs = 4096 FMAs.
// Used when the P accumulators of a warp (mp/64 x 8 tiles) fit in registers: T-mode, ncols <= 64, mp <= 128.
__device__ __forceinline__ void dmma884_p(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
constexpr int kP1LdY = kP1Cols + 4;
__host__ __device__ inline size_t psi1_tc_smem_bytes(int mp) { return ((size_t)kP1Rows * (mp + 4) + (size_t)kP1Rows * kP1LdY) * 8; }

template <int QP>
__global__ void __launch_bounds__(256, 2) psi1_fwd_tc_kernel(Psi1FwdParams p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double w1[kP1Rows][QP], mus[kP1Rows][QP], ld[kP1Rows][QP], lc[kP1Rows];
  const int LDM = p.mp + 4;
  double* tile = sm;                                  // [kP1Rows][LDM]   psi1
  double* ys = tile + kP1Rows * LDM;                  // [kP1Rows][kP1LdY] Y rows (row-major, zero padded)
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc4 = lane & 3;
  const int64_t items = p.nchunks * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  if (lo >= hi) return;
  const int mtw = (p.mp / 8 - warp + 7) / 8;          // m-tiles of this warp: warp, warp + 8 (mp <= 128)
  constexpr int CT = kP1Cols / 8;
  double pacc[2][CT][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) { pacc[i][j][0] = 0.0; pacc[i][j][1] = 0.0; }
  int cur_b = -1, seg = 0;
  double* mypart = nullptr;
  auto flush = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (i < mtw) {
#pragma unroll
        for (int j = 0; j < CT; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int m = (warp + 8 * i) * 8 + lr, c = j * 8 + 2 * lc4 + e;
            if (c < p.ncols) mypart[(size_t)m * p.cpad + c] = pacc[i][j][e];
            pacc[i][j][e] = 0.0;
          }
      }
    if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = cur_b;
    ++seg;
  };
  for (int64_t item = lo; item < hi; ++item) {
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * kP1Rows;
    const int nc = (int)min((int64_t)kP1Rows, p.n - n0);
    if (b != cur_b) {
      if (cur_b >= 0) flush();
      cur_b = b;
      mypart = p.part + ((size_t)blockIdx.x * p.nseg + seg) * p.mp * p.cpad;
    }
    __syncthreads();
    // Y rows of the item: asynchronous copy issued first, so that its DRAM latency overlaps the psi1 tile (exp) below
    for (int i = tid; i < kP1Rows * kP1Cols; i += T) {
      const int n = i / kP1Cols, c = i - n * kP1Cols;
      if (n < nc && c < p.ncols) cp_async8(ys + n * kP1LdY + c, p.y + (n0 + n) * p.d + c);
      else ys[n * kP1LdY + c] = 0.0;
    }
    cp_async_commit();
    psi1_row_terms<QP>(p.mu, p.s, p.gamma, p.alpha[b], n0, nc, p.q, b, w1, mus, ld, lc);
    {   // psi1 tile with leading dimension LDM (same arithmetic as psi1_tile)
      const int nparts = max(1, T / p.mp);
      for (int idx = tid; idx < p.mp * nparts; idx += T) {
        const int m = idx % p.mp, part = idx / p.mp;
        double zm[QP];
#pragma unroll
        for (int q = 0; q < QP; ++q) zm[q] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0;
        for (int n = part; n < kP1Rows; n += 2 * nparts) {
          const int n2 = n + nparts;
          double a0 = 0, a1 = 0;
#pragma unroll
          for (int q = 0; q < QP; ++q) {
            const double d0 = mus[n][q] - zm[q];
            a0 = fma(w1[n][q] * d0, d0, a0);
            if (n2 < kP1Rows) { const double d1 = mus[n2][q] - zm[q]; a1 = fma(w1[n2][q] * d1, d1, a1); }
          }
          tile[n * LDM + m] = (n < nc && m < p.m) ? exp_fast(fmax(fma(-0.5, a0, lc[n]), -1.0e8)) : 0.0;
          if (n2 < kP1Rows) tile[n2 * LDM + m] = (n2 < nc && m < p.m) ? exp_fast(fmax(fma(-0.5, a1, lc[n2]), -1.0e8)) : 0.0;
        }
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    if (p.psi1_out) {
      for (int i = tid; i < nc * p.m; i += T) {
        const int n = i / p.m, m = i % p.m;
        p.psi1_out[((int64_t)b * p.n + n0 + n) * p.m + m] = tile[n * LDM + m];
      }
    }
    // P[m][c] += sum_n psi1[n][m] Y[n][c]:  A[m][k = n] = tile[n][m],  B[k = n][c] = ys[n][c]
#pragma unroll
    for (int k0 = 0; k0 < kP1Rows; k0 += 4) {
      double bf[CT];
#pragma unroll
      for (int j = 0; j < CT; ++j) bf[j] = ys[(k0 + lc4) * kP1LdY + j * 8 + lr];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (i < mtw) {
          const double af = tile[(k0 + lc4) * LDM + (warp + 8 * i) * 8 + lr];
#pragma unroll
          for (int j = 0; j < CT; ++j) dmma884_p(pacc[i][j], af, bf[j]);
        }
    }
  }
  flush();
}

struct PReduceParams { const double* part; const int* tags; double* out; int grid, nseg, m, mp, ncols, cpad, b; int64_t nchunks; };
static __global__ void p_reduce_kernel(PReduceParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // (b, m, c)
  if (idx >= p.b * p.m * p.ncols) return;
  const int b = idx / (p.m * p.ncols), rem = idx % (p.m * p.ncols), m = rem / p.ncols, c = rem % p.ncols;
  int c_lo, c_hi; cta_range_of_cluster(b, p.nchunks, p.b, p.grid, c_lo, c_hi);
  double s = 0;
  for (int k = c_lo * p.nseg; k < (c_hi + 1) * p.nseg; ++k)
    if (p.tags[k] == b) s += p.part[((size_t)k * p.mp + m) * p.cpad + c];
  p.out[idx] = s;
}

// yy[d] = sum_n y_nd^2, kl = {sum mu^2, sum (s - log s)}: two-stage, fixed order.
struct ColSumParams { const double* mu; const double* s; const double* y; double* part; int64_t n; int d, q; };
static __global__ void __launch_bounds__(256) colsum_kernel(ColSumParams p) {
  __shared__ double red[32];
  const int64_t lo = p.n * blockIdx.x / gridDim.x, hi = p.n * (blockIdx.x + 1) / gridDim.x;
  double* out = p.part + (size_t)blockIdx.x * (p.d + 2);
  // yy: thread <-> column (strided), rows sequential: coalesced over d
  for (int d = threadIdx.x; d < p.d; d += blockDim.x) {
    double a = 0;
    for (int64_t n = lo; n < hi; ++n) { double v = p.y[n * p.d + d]; a = fma(v, v, a); }
    out[d] = a;
  }
  double a = 0, c = 0;
  for (int64_t i = lo * p.q + threadIdx.x; i < hi * p.q; i += blockDim.x) {
    double m_ = p.mu[i], sv = p.s[i];
    a = fma(m_, m_, a); c += sv - log(sv);
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0) out[p.d] = a;
  c = block_sum(c, red);
  if (threadIdx.x == 0) out[p.d + 1] = c;
}
static __global__ void colsum_reduce_kernel(const double* part, double* out, int nparts, int len) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = 0;
  for (int k = 0; k < nparts; ++k) s += part[(size_t)k * len + i];
  out[i] = s;
}

}  // namespace dpgp
