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
#include "common.cuh"

namespace dpgp {

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
