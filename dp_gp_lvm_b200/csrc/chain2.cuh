// Backward of the psi1 statistic fused with the chain of every per-(n,m) / per-(n,q) cotangent into the gradients
// of q(X) (mu, s), the inducing inputs Z and the kernel hyper-parameters (gamma, alpha).  Replaces TensorFlow
// autodiff through src/kernels/rbf_kernel.py:135-161 and through the N-contractions of
// src/models/dp_gp_lvm.py:132-145 / :638-658; supersedes g1_kernel + chain_bwd_kernel (psi1.cuh, chain.cuh), which
// round-tripped the [B,N,M] psi1 cotangent through HBM and did their contractions with scalar loops.
//
// Inputs per kernel-batch entry b (notation of chain.cuh):
//   a_nm = -1/2 dr_nm                      (dr: cotangent of r_nm from the fused psi2 backward)
//   b_nm = -1/2 psi1_nm (Y dP_b^T)_nm       (psi1 recomputed per 32-row tile, Y dP^T on the FP64 tensor cores)
//   dv_nq                                   (cotangent of v_nq from the fused psi2 backward)
// With delta = mu_nq - z_mq the chain needs, for x in {a, b},
//   per (n,q):  sum_m x delta,  sum_m x delta^2,  sum_m x        per (m,q):  sum_n x delta w~_nq   (w~ = w or w1)
// Expanding delta turns all of them into dense contractions, which run as FP64 DMMA (mma.sync m8n8k4):
//   [x]   (rows x M)  @ [1 | z | z^2]      (M x (1+2Q))   -> moments X0, X1_q, X2_q per row
//   [x]^T (M x rows)  @ [w~ | w~ mu]       (rows x 2Q)    -> V_mq, U_mq,   dz_mq = -2 (U_mq - z_mq V_mq)
// z and mu are centred by the column means of Z first, so the expansion loses no digits to |z|^2 >> delta^2.
// Fragment loads are bank-conflict free by construction: every shared matrix has a leading dimension = 4 or 12
// (mod 16) doubles.
#pragma once
#include "common.cuh"

namespace dpgp {

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

constexpr int kC2ColTile = 64;       // columns of Y per staged tile
constexpr int kC2MaxMT = 4;          // m-tiles (of 8) per warp: Mp <= 256

struct Chain2Params {
  const double* mu; const double* s; const double* y; const double* z; const double* gamma; const double* alpha;
  const double* dp;      // [B, M, ncols]   cotangent of P
  const double* dr;      // [B, N, mp]
  const double* dv;      // [B, N, QP]
  const double* dkl;     // [2]
  double* dmu; double* ds;           // [N,Q], complete on exit
  double* dzp;                       // [grid][mp*QP]    summed over b
  double* dgp;                       // [grid][B][QP]
  double* dap;                       // [grid][B]
  int64_t n; int d, q, m, mp, b, mode, ncols; int64_t nchunks;
  // Small N: the (chunk, b) items are also split over `bgroups` groups of clusters so that more than nchunks CTAs work;
  // group g then writes its share of dmu / ds into dmu_part / ds_part [bgroups][N][Q] (summed in fixed order by
  // chain2_rows_reduce_kernel).  bgroups == 1: results go straight to dmu / ds.  cgrid = CTAs per group.
  int bgroups, cgrid; double* dmu_part; double* ds_part;
};

template <int QP> __host__ __device__ constexpr int c2_jp() { return (1 + 2 * QP + 7) / 8 * 8; }
template <int QP> __host__ __device__ constexpr int c2_wp() { return (2 * QP + 7) / 8 * 8; }

template <int QP, int CR>
__host__ __device__ inline size_t chain2_smem_bytes(int mp) {
  const int LDM = mp + 4, LDY = kC2ColTile + 4, LDZ = c2_jp<QP>() + 4, LDW = c2_wp<QP>() + 4;
  // the final dz staging UO [mp][LDW] aliases BT | AT | Ys: 2 CR (mp + 4) + 68 CR >= mp LDW for every supported shape
  return (2 * (size_t)CR * LDM + (size_t)CR * LDY + (size_t)mp * LDZ + 2 * (size_t)CR * LDW + 2 * (size_t)CR * LDZ +
          5 * (size_t)CR * QP + CR + QP + 64) * 8;
}

template <int QP, int CR>
__global__ void __launch_bounds__(256, CR == 16 ? 2 : 1) psi1_bwd_chain_kernel(Chain2Params p) {
  extern __shared__ __align__(16) double sm[];
  constexpr int RT = CR / 8, JP = c2_jp<QP>(), JT = JP / 8, WP = c2_wp<QP>(), JT2 = WP / 8;
  constexpr int LDY = kC2ColTile + 4, LDZ = JP + 4, LDW = WP + 4, T = 256;
  constexpr int JS = 8 / (2 * RT);                      // splits of the j-tiles so that all 8 warps work in the moment GEMM
  const int LDM = p.mp + 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc4 = lane & 3;
  double* BT = sm;                                      // [CR][LDM]  psi1, then b
  double* AT = BT + (size_t)CR * LDM;                   // [CR][LDM]  a
  double* Ys = AT + (size_t)CR * LDM;                   // [CR][LDY]
  double* Zx = Ys + (size_t)CR * LDY;                   // [mp][LDZ]  [1 | zc | zc^2]
  double* Wa = Zx + (size_t)p.mp * LDZ;                 // [CR][LDW]  [w | w mu_c]
  double* Wb = Wa + (size_t)CR * LDW;                   // [CR][LDW]  [w1 | w1 mu_c]
  double* MOM = Wb + (size_t)CR * LDW;                  // [2][CR][LDZ]
  double* w_s = MOM + 2 * (size_t)CR * LDZ;             // [CR][QP] each
  double* w1_s = w_s + CR * QP;
  double* mu_s = w1_s + CR * QP;
  double* s_s = mu_s + CR * QP;
  double* dv_s = s_s + CR * QP;
  double* lc_s = dv_s + CR * QP;                        // [CR]
  double* zc = lc_s + CR;                               // [QP] column means of Z
  double* red = zc + QP;                                // [64]

  // ---- per-kernel constants: centred inducing inputs and their moment matrix
  if (tid < QP) {
    double a = 0;
    if (tid < p.q) for (int m = 0; m < p.m; ++m) a += p.z[m * p.q + tid];
    zc[tid] = a / p.m;
  }
  __syncthreads();
  for (int i = tid; i < p.mp * LDZ; i += T) {
    const int m = i / LDZ, j = i - m * LDZ;
    double v = 0.0;
    if (m < p.m) {
      if (j == 0) v = 1.0;
      else if (j <= QP) { const int q = j - 1; if (q < p.q) v = p.z[m * p.q + q] - zc[q]; }
      else if (j <= 2 * QP) { const int q = j - 1 - QP; if (q < p.q) { const double t = p.z[m * p.q + q] - zc[q]; v = t * t; } }
    }
    Zx[i] = v;
  }
  for (int i = tid; i < CR * LDW; i += T) { Wa[i] = 0.0; Wb[i] = 0.0; }
  const int ntw = (p.mp / 8 - warp + 7) / 8;            // m-tiles of this warp: warp, warp + 8, ...
  double U[kC2MaxMT][JT2][2];
#pragma unroll
  for (int i = 0; i < kC2MaxMT; ++i)
#pragma unroll
    for (int j = 0; j < JT2; ++j) { U[i][j][0] = 0.0; U[i][j][1] = 0.0; }
  const int TQ = (T / QP) * QP;
  const int nct = (p.ncols + kC2ColTile - 1) / kC2ColTile;
  const int npass = (p.mp / 8 + 15) / 16;               // passes of two m-tiles per warp in the Y dP^T product

  const int cta_c = blockIdx.x % p.cgrid, grp = blockIdx.x / p.cgrid;
  const int b_lo = (int)((int64_t)p.b * grp / p.bgroups), b_hi = (int)((int64_t)p.b * (grp + 1) / p.bgroups);
  double* out_mu = p.bgroups > 1 ? p.dmu_part + (size_t)grp * p.n * p.q : p.dmu;
  double* out_s = p.bgroups > 1 ? p.ds_part + (size_t)grp * p.n * p.q : p.ds;
  // The per-(n,q) inputs of the NEXT item (mu, s, dv) are loaded while the current one computes, and the running dmu / ds
  // sums of the current item at its start: their L2 / DRAM round trips were exposed at the head and the tail of every item.
  constexpr int NPT = (CR * QP + T - 1) / T;
  double pf_mu[NPT], pf_s[NPT], pf_dv[NPT];
  auto prefetch = [&](int pb, int64_t pck) {
    const int64_t pn0 = pck * CR;
    const int pnc = (int)min((int64_t)CR, p.n - pn0);
#pragma unroll
    for (int e = 0; e < NPT; ++e) {
      const int i = tid + e * T, n = i / QP, q = i - n * QP;
      const bool ok = i < CR * QP && n < pnc && q < p.q;
      pf_mu[e] = ok ? p.mu[(pn0 + n) * p.q + q] : 0.0;
      pf_s[e] = ok ? p.s[(pn0 + n) * p.q + q] : 1.0;
      pf_dv[e] = ok ? p.dv[((int64_t)pb * p.n + pn0 + n) * QP + q] : 0.0;
    }
  };
  if (b_lo < b_hi && cta_c < p.nchunks) prefetch(b_lo, cta_c);
  for (int b = b_lo; b < b_hi; ++b) {
    double dgam = 0.0, dalp = 0.0;
    const double alpha = p.alpha[b], lalpha = log(alpha);
    for (int64_t ck = cta_c; ck < p.nchunks; ck += p.cgrid) {
      const int64_t n0 = ck * CR;
      const int nc = (int)min((int64_t)CR, p.n - n0);
      double old_mu = 0.0, old_s = 0.0;
      if (NPT == 1 && b != b_lo && tid < CR * QP) {
        const int n = tid / QP, q = tid - n * QP;
        if (n < nc && q < p.q) { old_mu = out_mu[(n0 + n) * p.q + q]; old_s = out_s[(n0 + n) * p.q + q]; }
      }
      __syncthreads();
      // ---- dr tile: asynchronous copy straight into AT (raw dr; the factor -1/2 of a = -1/2 dr is applied where AT is
      //      read), issued first so that its DRAM latency overlaps S0 / S1 / S2 (it was a third of the stall samples)
      {
        const double* src = p.dr + ((int64_t)b * p.n + n0) * p.mp;
        const int vpr = p.mp / 2;                             // 16-byte vectors per row
        for (int i = tid; i < CR * vpr; i += T) {
          const int n = i / vpr, v = i - n * vpr;
          double* dst = AT + n * LDM + 2 * v;
          if (n < nc) cp_async16(dst, src + (size_t)n * p.mp + 2 * v);
          else { dst[0] = 0.0; dst[1] = 0.0; }
        }
        // first column tile of Y (T-mode): same treatment; later tiles (D > 64) are staged synchronously in S2
        if (p.mode != 1) {
          const int cw0 = min(kC2ColTile, p.ncols);
          for (int i = tid; i < CR * kC2ColTile; i += T) {
            const int n = i / kC2ColTile, c = i - n * kC2ColTile;
            if (n < nc && c < cw0) cp_async8(Ys + n * LDY + c, p.y + (n0 + n) * p.d + c);
            else Ys[n * LDY + c] = 0.0;
          }
        }
        cp_async_commit();
      }
      // ---- S0: per-(n,q) terms
#pragma unroll
      for (int e = 0; e < NPT; ++e) {
        const int i = tid + e * T;
        if (i >= CR * QP) break;
        const int n = i / QP, q = i - n * QP;
        double wv = 0, w1v = 0, mc = 0, sv = 1.0, dvv = 0, l1 = 0;
        if (n < nc && q < p.q) {
          const double g = p.gamma[b * p.q + q];
          sv = pf_s[e]; mc = pf_mu[e] - zc[q];
          const double den1 = fma(g, sv, 1.0);
          wv = g / fma(2.0 * g, sv, 1.0); w1v = g / den1; l1 = log(den1);
          dvv = pf_dv[e];
        }
        w_s[i] = wv; w1_s[i] = w1v; mu_s[i] = mc; s_s[i] = sv; dv_s[i] = dvv;
        Wa[n * LDW + q] = wv; Wa[n * LDW + QP + q] = wv * mc;
        Wb[n * LDW + q] = w1v; Wb[n * LDW + QP + q] = w1v * mc;
        MOM[i] = l1;                                    // scratch: log(g s + 1)
      }
      if (ck + p.cgrid < p.nchunks) prefetch(b, ck + p.cgrid);
      else if (b + 1 < b_hi) prefetch(b + 1, cta_c);
      __syncthreads();
      if (tid < CR) {
        double a = 0;
#pragma unroll
        for (int q = 0; q < QP; ++q) a += MOM[tid * QP + q];
        lc_s[tid] = lalpha - 0.5 * a;
      }
      __syncthreads();
      // ---- S1: psi1 tile (thread <-> (m, row phase), two rows in flight) and a = -1/2 dr
      {
        const int nparts = max(1, T / p.mp);
        for (int idx = tid; idx < p.mp * nparts; idx += T) {
          const int m = idx % p.mp, part = idx / p.mp;
          double zm[QP];
#pragma unroll
          for (int q = 0; q < QP; ++q) zm[q] = Zx[m * LDZ + 1 + q];
          for (int n = part; n < CR; n += 2 * nparts) {
            const int n2 = n + nparts;
            double a0 = 0, a1 = 0;
#pragma unroll
            for (int q = 0; q < QP; ++q) {
              const double d0 = mu_s[n * QP + q] - zm[q];
              a0 = fma(w1_s[n * QP + q] * d0, d0, a0);
              if (n2 < CR) { const double d1 = mu_s[n2 * QP + q] - zm[q]; a1 = fma(w1_s[n2 * QP + q] * d1, d1, a1); }
            }
            BT[n * LDM + m] = (n < nc && m < p.m) ? exp_fast(fmax(fma(-0.5, a0, lc_s[n]), -1.0e8)) : 0.0;
            if (n2 < CR) BT[n2 * LDM + m] = (n2 < nc && m < p.m) ? exp_fast(fmax(fma(-0.5, a1, lc_s[n2]), -1.0e8)) : 0.0;
          }
        }
      }
      // ---- S2: b = -1/2 psi1 o (Y dP^T)
      if (p.mode == 1) {
        __syncthreads();
        for (int i = tid; i < CR * p.mp; i += T) {
          const int n = i / p.mp, m = i - n * p.mp;
          if (n < nc && m < p.m) BT[n * LDM + m] *= -0.5 * p.y[(n0 + n) * p.d + b] * p.dp[(size_t)b * p.m + m];
        }
      } else {
        for (int pass = 0; pass < npass; ++pass) {
          const int mt0 = pass * 16 + warp, mt1 = mt0 + 8;            // this warp's two m-tiles of the pass
          const bool has0 = mt0 * 8 < p.mp, has1 = mt1 * 8 < p.mp;
          double C[RT][2][2];
#pragma unroll
          for (int r = 0; r < RT; ++r) { C[r][0][0] = C[r][0][1] = C[r][1][0] = C[r][1][1] = 0.0; }
          for (int ct = 0; ct < nct; ++ct) {
            const int cbase = ct * kC2ColTile, cw = min(kC2ColTile, p.ncols - cbase);
            if (ct == 0 && pass == 0) {
              cp_async_wait<0>();                           // the prefetched first tile (and the dr tile)
            } else {
              __syncthreads();
              for (int i = tid; i < CR * kC2ColTile; i += T) {
                const int n = i / kC2ColTile, c = i - n * kC2ColTile;
                Ys[n * LDY + c] = (n < nc && c < cw) ? p.y[(n0 + n) * p.d + cbase + c] : 0.0;
              }
            }
            __syncthreads();
            const int m0 = mt0 * 8 + lr, m1 = mt1 * 8 + lr;
            const double* dp0 = p.dp + ((size_t)b * p.m + (m0 < p.m ? m0 : 0)) * p.ncols + cbase;
            const double* dp1 = p.dp + ((size_t)b * p.m + (m1 < p.m ? m1 : 0)) * p.ncols + cbase;
            // the dP operands come from L2 (64 KB per cluster: no room in shared memory); four k-steps are loaded before
            // their DMMAs are issued so that one L2 round trip is shared by 8 loads (was: one per k-step, exposed)
            for (int k0 = 0; k0 < cw; k0 += 16) {
              double b0[4], b1[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int c = k0 + 4 * u + lc4;
                b0[u] = (has0 && m0 < p.m && c < cw) ? __ldg(dp0 + c) : 0.0;
                b1[u] = (has1 && m1 < p.m && c < cw) ? __ldg(dp1 + c) : 0.0;
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int c = k0 + 4 * u + lc4;
                if (k0 + 4 * u < cw) {
#pragma unroll
                  for (int r = 0; r < RT; ++r) {
                    const double a = Ys[(r * 8 + lr) * LDY + c];
                    dmma884(C[r][0], a, b0[u]);
                    dmma884(C[r][1], a, b1[u]);
                  }
                }
              }
            }
          }
#pragma unroll
          for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int mt = j ? mt1 : mt0;
              if (mt * 8 < p.mp) {
                double* dst = BT + (r * 8 + lr) * LDM + mt * 8 + 2 * lc4;
                dst[0] *= -0.5 * C[r][j][0]; dst[1] *= -0.5 * C[r][j][1];
              }
            }
        }
      }
      cp_async_wait<0>();
      __syncthreads();
      // ---- S3a: row moments  [a ; b] (rows x M) @ Zx (M x JP);  AT holds the raw dr: a = -1/2 dr
      {
        const int sel = warp / (RT * JS), rt = (warp / JS) % RT, js = warp % JS;
        const double ascale = sel ? 1.0 : -0.5;
        const double* src = (sel ? BT : AT) + (rt * 8 + lr) * LDM + lc4;
        double C[JT][2];
#pragma unroll
        for (int j = 0; j < JT; ++j) { C[j][0] = 0.0; C[j][1] = 0.0; }
#pragma unroll 4
        for (int k0 = 0; k0 < p.mp; k0 += 4) {
          const double a = ascale * src[k0];
          const double* zr = Zx + (k0 + lc4) * LDZ + lr;
#pragma unroll
          for (int j = 0; j < JT; ++j)
            if (j % JS == js) dmma884(C[j], a, zr[j * 8]);
        }
#pragma unroll
        for (int j = 0; j < JT; ++j)
          if (j % JS == js) {
            double* dst = MOM + ((size_t)sel * CR + rt * 8 + lr) * LDZ + j * 8 + 2 * lc4;
            dst[0] = C[j][0]; dst[1] = C[j][1];
          }
      }
      // ---- S3b: column side  a^T @ [w | w mu] + b^T @ [w1 | w1 mu], accumulated over chunks and clusters
#pragma unroll
      for (int k0 = 0; k0 < CR; k0 += 4) {
        double wa[JT2], wb[JT2];
#pragma unroll
        for (int j = 0; j < JT2; ++j) { wa[j] = Wa[(k0 + lc4) * LDW + j * 8 + lr]; wb[j] = Wb[(k0 + lc4) * LDW + j * 8 + lr]; }
#pragma unroll
        for (int i = 0; i < kC2MaxMT; ++i)
          if (i < ntw) {
            const int mcol = (warp + 8 * i) * 8 + lr;
            const double aa = -0.5 * AT[(k0 + lc4) * LDM + mcol], bb = BT[(k0 + lc4) * LDM + mcol];
#pragma unroll
            for (int j = 0; j < JT2; ++j) { dmma884(U[i][j], aa, wa[j]); dmma884(U[i][j], bb, wb[j]); }
          }
      }
      __syncthreads();
      // ---- S4: per-(n,q) results; stride TQ (a multiple of QP) keeps q = tid % QP fixed per thread
      for (int i = tid; tid < TQ && i < CR * QP; i += TQ) {
        const int n = i / QP, q = i - n * QP;
        if (n >= nc || q >= p.q) continue;
        const double* ma = MOM + (size_t)n * LDZ; const double* mb = MOM + ((size_t)CR + n) * LDZ;
        const double mc = mu_s[i];
        const double suma = ma[0], sumb = mb[0];
        const double sa1 = fma(mc, suma, -ma[1 + q]), sb1 = fma(mc, sumb, -mb[1 + q]);
        const double sa2 = fma(mc, fma(mc, suma, -2.0 * ma[1 + q]), ma[1 + QP + q]);
        const double sb2 = fma(mc, fma(mc, sumb, -2.0 * mb[1 + q]), mb[1 + QP + q]);
        const double g = p.gamma[b * p.q + q], sv = s_s[i], wv = w_s[i], w1v = w1_s[i];
        const double den = fma(2.0 * g, sv, 1.0), den1 = fma(g, sv, 1.0);
        const double dc = -suma, dlc = -2.0 * sumb, dvv = dv_s[i];
        double dmu = 2.0 * (wv * sa1 + w1v * sb1);
        double dsv = sa2 * (-2.0 * wv * wv) + dvv * (-0.5 * wv * wv) + dc * (-wv) + sb2 * (-w1v * w1v) + dlc * (-0.5 * w1v);
        dgam += sa2 / (den * den) + dvv * (-sv * g * den1 / (den * den)) + dc * (-sv / den) + sb2 / (den1 * den1) + dlc * (-0.5 * sv / den1);
        if (q == 0) dalp += (2.0 * dc + dlc) / alpha;
        const int64_t gi = (n0 + n) * p.q + q;
        if (b == 0) {                                   // the KL cotangents enter once, with cluster 0
          dmu += p.dkl[0] * 2.0 * (mc + zc[q]);
          dsv += p.dkl[1] * (1.0 - 1.0 / sv);
        }
        if (b == b_lo) { out_mu[gi] = dmu; out_s[gi] = dsv; }
        else if (NPT == 1) { out_mu[gi] = old_mu + dmu; out_s[gi] = old_s + dsv; }
        else { out_mu[gi] += dmu; out_s[gi] += dsv; }
      }
    }
    // ---- flush this cluster's dgamma / dalpha partials (fixed-order sums)
    __syncthreads();
    double* gq = AT;
    gq[tid] = (tid < TQ) ? dgam : 0.0;
    __syncthreads();
    if (tid < QP) {
      double a = 0;
      for (int j = tid; j < TQ; j += QP) a += gq[j];
      p.dgp[((size_t)blockIdx.x * p.b + b) * QP + tid] = a;
    }
    const double da = block_sum(dalp, red);
    if (tid == 0) p.dap[(size_t)blockIdx.x * p.b + b] = da;
  }
  // ---- flush dz: UO [mp][LDW] from the fragments, then dz_mq = -2 (U_mq - zc_mq V_mq)
  __syncthreads();
  double* UO = BT;                                      // [mp][LDW], aliases BT | AT | Ys (all dead now)
#pragma unroll
  for (int i = 0; i < kC2MaxMT; ++i)
    if (i < ntw) {
#pragma unroll
      for (int j = 0; j < JT2; ++j) {
        double* dst = UO + (size_t)((warp + 8 * i) * 8 + lr) * LDW + j * 8 + 2 * lc4;
        dst[0] = U[i][j][0]; dst[1] = U[i][j][1];
      }
    }
  __syncthreads();
  double* zp = p.dzp + (size_t)blockIdx.x * p.mp * QP;
  for (int i = tid; i < p.mp * QP; i += T) {
    const int m = i / QP, q = i - m * QP;
    zp[i] = -2.0 * (UO[m * LDW + QP + q] - Zx[m * LDZ + 1 + q] * UO[m * LDW + q]);
  }
}

// dmu / ds = sum over the cluster groups of the per-group partials (bgroups > 1 only), fixed order.
static __global__ void chain2_rows_reduce_kernel(const double* mu_part, const double* s_part, double* dmu, double* ds,
                                                 int64_t len, int groups) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    double a = 0, c = 0;
    for (int g = 0; g < groups; ++g) { a += mu_part[(size_t)g * len + i]; c += s_part[(size_t)g * len + i]; }
    dmu[i] = a; ds[i] = c;
  }
}

// Final fixed-order sums of the per-CTA partials of psi1_bwd_chain_kernel (+ the dD chain of zchain_kernel).
static __global__ void chain2_reduce_kernel(const double* dzp, const double* dgp, const double* dap, const double* dzd,
                                            double* dz, double* dgamma, double* dalpha, int grid, int b_count, int m, int mp,
                                            int q, int qp) {
  // one warp per output, lanes stride the per-CTA partial slots, fixed-order shuffle tree (a thread per output walked
  // the <= 296 slots serially: 58 us of L2 latency per evaluation, 5 % of a training iteration at the reference's sizes)
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nz = m * q, ng = b_count * q;
  if (i >= nz + ng + b_count) return;
  double s = 0;
  if (i < nz) {
    const int mm_ = i / q, qq = i % q;
    for (int c = lane; c < grid; c += 32) s += dzp[(size_t)c * mp * qp + mm_ * qp + qq];
    for (int b = lane; b < b_count; b += 32) s += dzd[(size_t)b * nz + i];
    s = warp_sum(s);
    if (lane == 0) dz[i] = s;
  } else if (i < nz + ng) {
    const int j = i - nz, b = j / q, qq = j % q;
    for (int c = lane; c < grid; c += 32) s += dgp[((size_t)c * b_count + b) * qp + qq];
    s = warp_sum(s);
    if (lane == 0) dgamma[j] = s;
  } else {
    const int b = i - nz - ng;
    for (int c = lane; c < grid; c += 32) s += dap[(size_t)c * b_count + b];
    s = warp_sum(s);
    if (lane == 0) dalpha[b] = s;
  }
}

}  // namespace dpgp
