// Chain of the per-(n,m) and per-(n,q) cotangents produced by the psi1/psi2 backward kernels into the
// gradients of q(X) (mu, s), the inducing inputs Z and the kernel hyper-parameters (gamma, alpha).
//
// Inputs per kernel-batch entry b:
//   dr  [N,mp]  cotangent of r_nm   = 1/2 c_n - 1/2 sum_q w_nq (mu_nq - z_mq)^2        (psi2, psi2_bwd_n_kernel)
//   dv  [N,QP]  cotangent of v_nq   = -1/2 g^2 s/(2 g s + 1)                            (psi2)
//   bco [N,mp]  = -1/2 * cotangent of log psi1_nm, log psi1_nm = lc_n - 1/2 sum_q w1_nq (mu_nq - z_mq)^2   (g1_kernel)
// With a_nm = -1/2 dr_nm, b_nm = bco_nm, delta = mu_nq - z_mq:
//   dmu_nq  = 2 (w sum_m a delta + w1 sum_m b delta)          dz_mq = -2 sum_n delta (a w + b w1)
//   dw_nq   = sum_m a delta^2,  dw1_nq = sum_m b delta^2,  dc_n = -sum_m a,  dlc_n = -2 sum_m b
// and w = g/(2gs+1), w1 = g/(gs+1), c_n = 2 log alpha - 1/2 sum_q log(2gs+1), lc_n = log alpha - 1/2 sum_q log(gs+1)
// are chained into s, gamma, alpha (derivatives written out in the kernel).  The KL cotangents
// (dkl0 * 2 mu, dkl1 * (1 - 1/s)) initialise dmu / ds on the first b.
#pragma once
#include "../common.cuh"
#include "../psi1.cuh"

namespace dpgp {

// ----------------------------------------------------------------------------------- psi1 backward, part 1
// bco[b,n,m] = -1/2 psi1_nm * sum_c Y[n, col(b,c)] dP[b,m,c]   (cotangent of sum_q w1 (mu - z)^2 ... see chain.cuh)
struct G1Params {
  const double* mu; const double* s; const double* y; const double* z; const double* gamma; const double* alpha;
  const double* dp;      // [B, M, ncols]
  double* bco;           // [B, N, mp]
  int64_t n; int d, q, m, mp, b, mode, ncols; int64_t nchunks;
};

template <int QP>
__global__ void __launch_bounds__(256) g1_kernel(G1Params p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double w1[kP1Rows][QP], mus[kP1Rows][QP], ld[kP1Rows][QP], lc[kP1Rows];
  double* tile = sm;                                  // psi1 [kP1Rows][mp]
  double* yt = tile + kP1Rows * p.mp;                 // [kP1Cols][kP1Rows]
  double* dpt = yt + kP1Cols * kP1Rows;               // [kP1Cols][mp]
  const int tid = threadIdx.x, T = blockDim.x;
  const int64_t items = p.nchunks * p.b;
  const int mtiles = p.mp / 4, ntl = kP1Rows / 4;
  const int nct = (p.ncols + kP1Cols - 1) / kP1Cols;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * kP1Rows;
    const int nc = (int)min((int64_t)kP1Rows, p.n - n0);
    __syncthreads();
    psi1_row_terms<QP>(p.mu, p.s, p.gamma, p.alpha[b], n0, nc, p.q, b, w1, mus, ld, lc);
    psi1_tile<QP>(p.z, p.m, p.mp, p.q, nc, w1, mus, lc, tile);
    const int col0 = (p.mode == 1) ? b : 0;
    // up to (mtiles * ntl) / T register tiles of 4 rows x 4 inducing points per thread (<= 2 at M = 256)
    double acc[2][4][4];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][i][j] = 0.0;
    for (int ct = 0; ct < nct; ++ct) {
      const int cbase = ct * kP1Cols;
      const int cw = min(kP1Cols, p.ncols - cbase);
      __syncthreads();
      for (int i = tid; i < kP1Cols * kP1Rows; i += T) {
        int n = i / kP1Cols, c = i % kP1Cols;
        double v = 0.0;
        if (n < nc && c < cw) v = p.y[(n0 + n) * p.d + col0 + cbase + c];
        yt[c * kP1Rows + n] = v;
      }
      for (int i = tid; i < kP1Cols * p.mp; i += T) {
        int m = i / kP1Cols, c = i % kP1Cols;
        double v = 0.0;
        if (m < p.m && c < cw) v = p.dp[((size_t)b * p.m + m) * p.ncols + cbase + c];
        dpt[c * p.mp + m] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int t = tid + k * T;
        if (t >= mtiles * ntl) continue;
        const int m0 = (t % mtiles) * 4, r0 = (t / mtiles) * 4;
        for (int c = 0; c < cw; ++c) {
          const double2 y01 = *reinterpret_cast<const double2*>(yt + c * kP1Rows + r0);
          const double2 y23 = *reinterpret_cast<const double2*>(yt + c * kP1Rows + r0 + 2);
          const double2 d01 = *reinterpret_cast<const double2*>(dpt + c * p.mp + m0);
          const double2 d23 = *reinterpret_cast<const double2*>(dpt + c * p.mp + m0 + 2);
          const double yv[4] = {y01.x, y01.y, y23.x, y23.y};
          const double dv[4] = {d01.x, d01.y, d23.x, d23.y};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[k][i][j] = fma(yv[i], dv[j], acc[k][i][j]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int t = tid + k * T;
      if (t >= mtiles * ntl) continue;
      const int m0 = (t % mtiles) * 4, r0 = (t / mtiles) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (r0 + i >= nc) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          p.bco[((int64_t)b * p.n + n0 + r0 + i) * p.mp + m0 + j] = -0.5 * acc[k][i][j] * tile[(r0 + i) * p.mp + m0 + j];
      }
    }
  }
}


constexpr int kChRows = 32;

struct ChainParams {
  const double* mu; const double* s; const double* z; const double* gamma; const double* alpha;
  const double* dr; const double* dv; const double* bco; const double* dkl;
  double* dmu; double* ds;           // [N,Q], complete on exit
  double* dzp;                       // [grid][B][mp*QP]
  double* dgp;                       // [grid][B][QP]
  double* dap;                       // [grid][B]
  int64_t n; int q, m, mp, b; int64_t nchunks;
};

template <int QP>
__global__ void __launch_bounds__(256) chain_bwd_kernel(ChainParams p) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double w[kChRows][QP], w1[kChRows][QP], mus[kChRows][QP], ss[kChRows][QP], dvs[kChRows][QP];
  __shared__ double red[32];
  double* at = sm;                         // [kChRows][mp]
  double* bt = at + kChRows * p.mp;        // [kChRows][mp]
  double* zs = bt + kChRows * p.mp;        // [mp][QP]
  const int tid = threadIdx.x, T = blockDim.x;
  constexpr int KZ = (kMaxM * kMaxQ) / 256;          // dz accumulators per thread, worst case
  for (int i = tid; i < p.mp * QP; i += T) { int m = i / QP, q = i % QP; zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0; }
  const int nzq = p.mp * QP;
  const int TQ = (T / QP) * QP;
  for (int b = 0; b < p.b; ++b) {
    double dz[KZ];
#pragma unroll
    for (int k = 0; k < KZ; ++k) dz[k] = 0.0;
    double dgam = 0.0, dalp = 0.0;           // per-thread partials; thread (n,q) contributes to dgamma_q with q = idx % QP
    const double alpha = p.alpha[b];
    for (int64_t ck = blockIdx.x; ck < p.nchunks; ck += gridDim.x) {
      const int64_t n0 = ck * kChRows;
      const int nc = (int)min((int64_t)kChRows, p.n - n0);
      __syncthreads();
      for (int i = tid; i < kChRows * p.mp; i += T) {
        const int n = i / p.mp;
        double a = 0, bb = 0;
        if (n < nc) {
          const int64_t g = ((int64_t)b * p.n + n0) * p.mp + i;
          a = -0.5 * p.dr[g]; bb = p.bco[g];
        }
        at[i] = a; bt[i] = bb;
      }
      for (int i = tid; i < kChRows * QP; i += T) {
        const int n = i / QP, q = i % QP;
        double wv = 0, w1v = 0, m_ = 0, sv = 1.0, dvv = 0;
        if (n < nc && q < p.q) {
          const double g = p.gamma[b * p.q + q];
          sv = p.s[(n0 + n) * p.q + q]; m_ = p.mu[(n0 + n) * p.q + q];
          wv = g / fma(2.0 * g, sv, 1.0); w1v = g / fma(g, sv, 1.0);
          dvv = p.dv[((int64_t)b * p.n + n0 + n) * QP + q];
        }
        w[n][q] = wv; w1[n][q] = w1v; mus[n][q] = m_; ss[n][q] = sv; dvs[n][q] = dvv;
      }
      __syncthreads();
      // ---- n side: thread <-> (n, q); stride TQ (a multiple of QP) keeps q = tid % QP fixed per thread
      for (int i = tid; tid < TQ && i < kChRows * QP; i += TQ) {
        const int n = i / QP, q = i % QP;
        if (n >= nc || q >= p.q) continue;
        const double m_ = mus[n][q];
        double sa1 = 0, sa2 = 0, sb1 = 0, sb2 = 0, suma = 0, sumb = 0;
        for (int m = 0; m < p.m; ++m) {
          const double d = m_ - zs[m * QP + q];
          const double a = at[n * p.mp + m], bb = bt[n * p.mp + m];
          const double ad = a * d, bd = bb * d;
          sa1 += ad; sa2 = fma(ad, d, sa2); sb1 += bd; sb2 = fma(bd, d, sb2); suma += a; sumb += bb;
        }
        const double g = p.gamma[b * p.q + q], sv = ss[n][q], wv = w[n][q], w1v = w1[n][q];
        const double den = fma(2.0 * g, sv, 1.0), den1 = fma(g, sv, 1.0);
        const double dc = -suma, dlc = -2.0 * sumb, dvv = dvs[n][q];
        double dmu = 2.0 * (wv * sa1 + w1v * sb1);
        double dsv = sa2 * (-2.0 * wv * wv) + dvv * (-0.5 * wv * wv) + dc * (-wv) + sb2 * (-w1v * w1v) + dlc * (-0.5 * w1v);
        dgam += sa2 / (den * den) + dvv * (-sv * g * den1 / (den * den)) + dc * (-sv / den) + sb2 / (den1 * den1) + dlc * (-0.5 * sv / den1);
        if (q == 0) dalp += (2.0 * dc + dlc) / alpha;
        const int64_t gi = (n0 + n) * p.q + q;
        if (b == 0) {
          dmu += p.dkl[0] * 2.0 * m_;
          dsv += p.dkl[1] * (1.0 - 1.0 / sv);
          p.dmu[gi] = dmu; p.ds[gi] = dsv;
        } else {
          p.dmu[gi] += dmu; p.ds[gi] += dsv;
        }
      }
      // ---- m side: thread <-> (m, q), accumulators persist over this CTA's chunks
#pragma unroll
      for (int k = 0; k < KZ; ++k) {
        const int i = tid + k * T;
        if (i < nzq) {
          const int m = i / QP, q = i % QP;
          const double zv = zs[i];
          double acc = 0;
          for (int n = 0; n < nc; ++n) {
            const double d = mus[n][q] - zv;
            acc = fma(d, fma(at[n * p.mp + m], w[n][q], bt[n * p.mp + m] * w1[n][q]), acc);
          }
          dz[k] = fma(-2.0, acc, dz[k]);
        }
      }
    }
    // flush this b
    double* zp = p.dzp + ((size_t)blockIdx.x * p.b + b) * nzq;
#pragma unroll
    for (int k = 0; k < KZ; ++k) { const int i = tid + k * T; if (i < nzq) zp[i] = dz[k]; }
    // dgamma_q: threads tid < TQ hold the partial of q = tid % QP; fixed-order sum per q
    __syncthreads();
    double* gq = at;                         // reuse the tile: [T]
    gq[tid] = (tid < TQ) ? dgam : 0.0;
    __syncthreads();
    if (tid < QP) {
      double a = 0;
      for (int j = tid; j < TQ; j += QP) a += gq[j];
      p.dgp[((size_t)blockIdx.x * p.b + b) * QP + tid] = a;
    }
    double da = block_sum(dalp, red);
    if (tid == 0) p.dap[(size_t)blockIdx.x * p.b + b] = da;
  }
}

}  // namespace dpgp
