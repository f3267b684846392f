// Fused backward of the psi2 statistic, second generation (bwd_variant 8): the two N / pair contractions of
// psi2_bwd_fused_kernel run as FP64 DMMA (mma.sync.m8n8k4) on the first 8 latent dimensions and as DFMA on the rest.
//
//   g_np  = Gs_p exp(r_nm + r_nm' + sum_q v_nq D_pq)      dr_nm = sum_{m'} g_n(m,m')
//   dv_nq = sum_p g_np D_pq                                 dD_pq = sum_n g_np v_nq          (psi2_bwd_fused.cuh)
//
// Why.  psi2_bwd_fused_kernel is bound by its warp-instruction count (80.5 per unit at Q = 10, issue 48 %, FP64 pipe 55 %:
// profiles/r01_fused_final.md, profiles/r02_umma.md 4.1), and 30 of its instructions per unit are the two contractions written as
// DFMAs with their operand loads.  DMMA and DFMA share one pipe on B200, so a DMMA buys no FLOPs -- but one m8n8k4 issues 256
// FMAs where a DFMA issues 32.  Round 1's tensor-core attempt put the EXPONENT on DMMA and lost to padding (Q = 10 -> 12, 16);
// here the split is exact: q 0..7 fill the n = 8 side of the MMA completely, q 8.. stay on DFMA, so the pipe load is unchanged
// (92 cycles per unit) while the instruction count drops by ~14 per unit.
//
// Per warp and 16-pair step (two block rows of the warp's 8 x 8 pair block, 64 rows):
//   phase 1   lane <-> rows (lane, lane + 32): exponent, exp, row / column sums of g, g -> gt[16 pairs][64 rows, swizzled]; dv only for q >= 8
//   phase 2a  dD[pair][q < 8] += gt[pair][row] v[row][q]:  A = gt (m = pairs, k = rows), B = vt, 2 m-tiles x 16 k-steps
//   phase 2b  dv[row][q < 8]  += gt[pair][row] D[pair][q]:  A = gt read transposed (m = rows, k = pairs), B = pair table,
//             8 m-tiles x 4 k-steps; the accumulators (16 doubles) live in registers for the whole item
//   phase 2c  dD[pair][q >= 8] by DFMA, lane <-> (pair, half of the rows), rows rotated per pair group (conflict-free)
// Leading dimensions 12 (vt), 12 / 20 (pair table) are = 4 or 12 mod 16 and the g tile is XOR-swizzled (row ^ 4 (pair & 3), no
// padding: shared memory is full), which makes every fragment load of the three phases conflict-free.
// The dD totals are contracted with 2 (z_m - z_m') and added into per-warp dz slices exactly as in bwd_variant 6 (the C-fragment
// layout needs 17 shuffles per step where the DFMA layout needed 52).  Results are bitwise reproducible.
#pragma once
#include "psi2_bwd_fused.cuh"
#include "chain2.cuh"          // dmma884

namespace dpgp {

constexpr int kMmaGS = 64, kMmaLDV = 12;          // gt rows are XOR-swizzled (row ^ 4 (pair & 3)) instead of padded
template <int QP> __host__ __device__ constexpr int mma_ds() { return QP <= 10 ? 12 : 20; }

// smem (doubles): rT[mp*65] | drT[mp*65] | zs[mp*QP] | etab[256] | vt[64][12] | dtab[8][16*DS] | gt[8][16*64] (aliases xdv[8][QP][65])
template <int QP>
__host__ __device__ inline size_t mma_smem_bytes(int mp) {
  const size_t gt = (size_t)kFusedWarps * kFusedPB * kMmaGS, xd = (size_t)kFusedWarps * QP * 65;
  return (2 * (size_t)mp * 65 + (size_t)mp * QP + kExpTabSize + 64 * kMmaLDV + (size_t)kFusedWarps * kFusedPB * mma_ds<QP>() + (gt > xd ? gt : xd)) * 8;
}

template <int QP, int EXPV>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) psi2_bwd_mma_kernel(Psi2BwdFusedParams p) {
  static_assert(QP >= 8 && QP <= 12, "psi2_bwd_mma_kernel: 8 <= QP <= 12");
  extern __shared__ __align__(16) double sm[];
  constexpr int R = 2, RS = 65, ROWS = 64, DS = mma_ds<QP>(), T = kFusedWarps * 32, PB = kFusedPB, KU = 4;
  constexpr int GS = kMmaGS, LDV = kMmaLDV, QR = QP - 8;              // QR latent dimensions stay on DFMA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;                            // fragment coordinates
  double* rT = sm;
  double* drT = rT + (size_t)p.mp * RS;
  double* zs = drT + (size_t)p.mp * RS;
  double* etab = zs + (size_t)p.mp * QP;
  double* vt = etab + kExpTabSize;                      // [ROWS][LDV]
  double* dtab = vt + (size_t)ROWS * LDV;
  double* gtab = dtab + (size_t)kFusedWarps * PB * DS;
  double* xdv = gtab;                                   // alias, used only between the last round and the next fill
  double* dtw = dtab + (size_t)warp * PB * DS;
  double* gtw = gtab + (size_t)warp * PB * GS;

  for (int i = tid; i < p.mp * QP; i += T) { const int m = i / QP, q = i % QP; zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0; }
  load_exp_table(etab, p.exptab);
  Exp<EXPV> ex; ex.init(etab);
  const uint64_t keep = l2_evict_last_policy();
  double* dzr = p.part + ((size_t)blockIdx.x * kFusedWarps + warp) * 2 * p.mp * QP;      // row-side slice of this warp
  double* dzc = dzr + (size_t)p.mp * QP;                                                 // column side
  const int p2_pair = lane >> 1, p2_qh = lane & 1;      // pair table build: lane <-> (pair of the step, q half)
  constexpr int QH = QP / 2;
  const int c_p = lane & 15, c_rh = lane >> 4;          // phase 2c: lane <-> (pair, half of the rows)

  const int64_t items = p.ngroups * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  for (int64_t item = lo; item < hi; ++item) {
    const int b = (int)(item / p.ngroups);
    const int64_t n0 = (item % p.ngroups) * ROWS;
    const int nc = (int)min((int64_t)ROWS, p.n - n0);
    __syncthreads();                                    // previous group's drain has finished with rT / drT / xdv
    {
      const double* src = p.r + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < ROWS * p.mp; idx += T) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        rT[(size_t)m * RS + row] = (row < nc) ? __ldcs(src + idx) : kRClamp;
      }
      for (int idx = tid; idx < p.mp * RS; idx += T) drT[idx] = 0.0;
      const double* vsrc = p.v + ((int64_t)b * p.n + n0) * QP;
      for (int idx = tid; idx < ROWS * LDV; idx += T) {
        const int row = idx / LDV, q = idx - row * LDV;
        vt[idx] = (row < nc && q < QP) ? __ldcs(vsrc + row * QP + q) : 0.0;
      }
    }
    double vq[R][QP], dvr[R][QR > 0 ? QR : 1], cdv[8][2];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int row = lane + 32 * rr;
      const double* vs = p.v + ((int64_t)b * p.n + n0 + (row < nc ? row : 0)) * QP;
#pragma unroll
      for (int q = 0; q < QP; q += 2) {
        const double2 t2 = __ldcs(reinterpret_cast<const double2*>(vs + q));
        vq[rr][q] = (row < nc) ? t2.x : 0.0; vq[rr][q + 1] = (row < nc) ? t2.y : 0.0;
      }
#pragma unroll
      for (int q = 0; q < (QR > 0 ? QR : 1); ++q) dvr[rr][q] = 0.0;
    }
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) { cdv[mt][0] = 0.0; cdv[mt][1] = 0.0; }
    const double* gb = p.gbar + (size_t)b * p.m * p.m;
    auto load_w = [&](unsigned short it, double (&w)[2]) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int id = lane + 32 * e;
        w[e] = (it == kSchedIdle) ? 0.0 : sym_cotangent(gb, 8 * (it >> 8) + (id >> 3), 8 * (it & 255) + (id & 7), p.m);
      }
    };
    double wc[2], wn[2] = {0.0, 0.0};
    load_w(0 < p.nrounds ? p.sched[warp] : kSchedIdle, wc);
    __syncthreads();

    for (int round = 0; round < p.nrounds; ++round) {
      const unsigned short it = p.sched[round * kFusedWarps + warp];
      if (round + 1 < p.nrounds) load_w(p.sched[(round + 1) * kFusedWarps + warp], wn);
      if (it != kSchedIdle) {
        const int bi = it >> 8, bj = it & 255;
        const double* rcol = rT + (size_t)(8 * bj) * RS + lane;
        double cs[8][R];
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int rr = 0; rr < R; ++rr) cs[k][rr] = 0.0;
#pragma unroll 1
        for (int half = 0; half < 64 / PB; ++half) {
          // ---- table of this step's 16 pairs: D[pair][q] and the symmetrised cotangent; lane <-> (pair, q half)
          {
            const int i = 2 * half + (p2_pair >> 3), k = p2_pair & 7, m = 8 * bi + i, c = 8 * bj + k;
#pragma unroll
            for (int j = 0; j < QH; ++j) {
              const int q = p2_qh * QH + j;
              const double d = zs[m * QP + q] - zs[c * QP + q];
              dtw[p2_pair * DS + q] = d * d;
            }
            const double wv = __shfl_sync(0xffffffffu, (half & 2) ? wc[1] : wc[0], 16 * (half & 1) + p2_pair);
            if (p2_qh == 0) dtw[p2_pair * DS + QP] = wv;
          }
          __syncwarp();
          // ---- phase 1: lane <-> rows
#pragma unroll 1
          for (int i2 = 0; i2 < 2; ++i2) {
            const int i = 2 * half + i2;
            double rm[R], rs[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { rm[rr] = rT[(size_t)(8 * bi + i) * RS + lane + 32 * rr]; rs[rr] = 0.0; }
#pragma unroll
            for (int k0 = 0; k0 < 8; k0 += KU) {
              double dq[KU][QP], e[KU * R], w[KU * R], g[KU * R];
#pragma unroll
              for (int u = 0; u < KU; ++u) {
                const double* dt = dtw + (i2 * 8 + k0 + u) * DS;
#pragma unroll
                for (int q = 0; q < QP; q += 2) { const double2 t2 = *reinterpret_cast<const double2*>(dt + q); dq[u][q] = t2.x; dq[u][q + 1] = t2.y; }
                const double wgt = dt[QP];
                double ea[R], eb[R];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) { ea[rr] = rm[rr]; eb[rr] = rcol[(size_t)(k0 + u) * RS + 32 * rr]; w[u * R + rr] = wgt; }
#pragma unroll
                for (int q = 0; q < QP; q += 2)
#pragma unroll
                  for (int rr = 0; rr < R; ++rr) { ea[rr] = fma(vq[rr][q], dq[u][q], ea[rr]); eb[rr] = fma(vq[rr][q + 1], dq[u][q + 1], eb[rr]); }
#pragma unroll
                for (int rr = 0; rr < R; ++rr) e[u * R + rr] = ea[rr] + eb[rr];
              }
              exp_scaled_k<EXPV, KU * R>(ex, e, w, g);
#pragma unroll
              for (int u = 0; u < KU; ++u) {
                double* gdst = gtw + (size_t)(i2 * 8 + k0 + u) * GS + (lane ^ (4 * ((k0 + u) & 3)));
#pragma unroll
                for (int rr = 0; rr < R; ++rr) gdst[32 * rr] = g[u * R + rr];
#pragma unroll
                for (int rr = 0; rr < R; ++rr) { rs[rr] += g[u * R + rr]; cs[k0 + u][rr] += g[u * R + rr]; }
#pragma unroll
                for (int q = 0; q < QR; ++q)
#pragma unroll
                  for (int rr = 0; rr < R; ++rr) dvr[rr][q] = fma(g[u * R + rr], dq[u][8 + q], dvr[rr][q]);
              }
            }
#pragma unroll
            for (int rr = 0; rr < R; ++rr) drT[(size_t)(8 * bi + i) * RS + lane + 32 * rr] += rs[rr];
          }
          __syncwarp();
          // ---- phase 2a: dD[pair][q < 8] on the tensor cores; C fragment: pair = 8 mt + lr, q = 2 lc + {0, 1}
          double cdd[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
          {
            const double* a0 = gtw + (size_t)lr * GS + lc;
            const double* a1 = a0 + 8 * GS;
            const double* bv = vt + (size_t)lc * LDV + lr;
            const int sw = lr & 3;                                  // swizzle of this lane's pair rows: row 4 ks + lc sits at 4 (ks ^ sw) + lc
            // (four independent accumulator pairs with batched fragment loads, and the same for phase 2c, measured 102.2 ms
            // against 95.9 ms for these plain loops at 262 144 rows: ptxas' schedule of the 254-register body, not the chains, decides)
#pragma unroll
            for (int ks = 0; ks < ROWS / 4; ++ks) {
              const double bb = bv[(size_t)ks * 4 * LDV];
              dmma884(cdd[0], a0[4 * (ks ^ sw)], bb);
              dmma884(cdd[1], a1[4 * (ks ^ sw)], bb);
            }
          }
          // ---- phase 2b: dv[row][q < 8]; A = gt read transposed (row = 8 mt + lr, pair = 4 ks + lc), B = pair table
          {
#pragma unroll
            for (int ks = 0; ks < PB / 4; ++ks) {
              const double bb = dtw[(ks * 4 + lc) * DS + lr];
              // pair & 3 == lc: row 8 mt + lr sits at (8 mt + lr) ^ 4 lc = 8 (mt ^ (lc >> 1)) + (lr ^ 4 (lc & 1))
              const double* ar = gtw + (size_t)(ks * 4 + lc) * GS + (lr ^ (4 * (lc & 1)));
              const int x8 = 8 * (lc >> 1);
              const double* pe = ar + x8;
              const double* po = ar - x8;
#pragma unroll
              for (int mt = 0; mt < 8; ++mt) dmma884(cdv[mt], (mt & 1) ? po[8 * mt] : pe[8 * mt], bb);
            }
          }
          // ---- phase 2c: dD[pair][q >= 8] by DFMA; lane <-> (pair c_p, row half c_rh); rows rotated by the pair group (with the swizzle: conflict-free)
          double t8[QR > 0 ? QR : 1];
          if constexpr (QR > 0) {
            double acc[QR];
#pragma unroll
            for (int q = 0; q < QR; ++q) acc[q] = 0.0;
            const double* gp = gtw + (size_t)c_p * GS + c_rh * 32;
            const double* vp = vt + (size_t)(c_rh * 32) * LDV + 8;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int row = (i + (c_p >> 2)) & 31;
              const double gg = gp[row ^ (4 * (c_p & 3))];
              if constexpr (QR == 2) {
                const double2 v2 = *reinterpret_cast<const double2*>(vp + (size_t)row * LDV);
                acc[0] = fma(gg, v2.x, acc[0]); acc[1] = fma(gg, v2.y, acc[1]);
              } else {
#pragma unroll
                for (int q = 0; q < QR; ++q) acc[q] = fma(gg, vp[(size_t)row * LDV + q], acc[q]);
              }
            }
            // merge the two row halves
#pragma unroll
            for (int q = 0; q < QR; ++q) t8[q] = acc[q] + __shfl_xor_sync(0xffffffffu, acc[q], 16);
          }
          // ---- dz folding: t = (z_m - z_m') dD (the factor 2 is applied by dz_fused_reduce_kernel)
          {
            // 2a values: pair (i2 = mt, k = lr), q = 2 lc + j
            double ta[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const int mrow = 8 * bi + 2 * half + mt, mcol = 8 * bj + lr;
#pragma unroll
              for (int j = 0; j < 2; ++j) ta[mt][j] = (zs[(size_t)mrow * QP + 2 * lc + j] - zs[(size_t)mcol * QP + 2 * lc + j]) * cdd[mt][j];
            }
            // 2c values: pair c_p (i2 = c_p >> 3, k = c_p & 7); both row-half lanes hold the total: lane half h keeps q = 8 + h (QR <= 2: one q
            // per half; QR = 4: two per half)
            constexpr int QK = QR > 0 ? (QR + 1) / 2 : 1;
            double tc[QK];
            if constexpr (QR > 0) {
              const int mrow = 8 * bi + 2 * half + (c_p >> 3), mcol = 8 * bj + (c_p & 7);
#pragma unroll
              for (int e = 0; e < QK; ++e) {
                const int q = 8 + c_rh * QK + e;
                const double dd = (QR == 1) ? t8[0] : (c_rh ? t8[QK + e < QR ? QK + e : QR - 1] : t8[e]);
                tc[e] = (q < QP) ? (zs[(size_t)mrow * QP + q] - zs[(size_t)mcol * QP + q]) * dd : 0.0;
              }
            }
            // all shuffles first, in converged code
            double ra[2][2], rc[QK], cc[QK];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                double a = ta[mt][j] + __shfl_xor_sync(0xffffffffu, ta[mt][j], 4);       // over the 8 columns (lr) of the block row
                a += __shfl_xor_sync(0xffffffffu, a, 8);
                ra[mt][j] = a + __shfl_xor_sync(0xffffffffu, a, 16);
              }
            if constexpr (QR > 0) {
#pragma unroll
              for (int e = 0; e < QK; ++e) {
                double a = tc[e] + __shfl_xor_sync(0xffffffffu, tc[e], 1);               // over the 8 columns (c_p & 7)
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                rc[e] = a + __shfl_xor_sync(0xffffffffu, a, 4);
                cc[e] = tc[e] + __shfl_xor_sync(0xffffffffu, tc[e], 8);                 // over the two block rows
              }
            }
            if (lr == 0) {
#pragma unroll
              for (int mt = 0; mt < 2; ++mt) {
                double* dst = dzr + (size_t)(8 * bi + 2 * half + mt) * QP + 2 * lc;
                red_add_f64_keep(dst, ra[mt][0], keep); red_add_f64_keep(dst + 1, ra[mt][1], keep);
              }
            }
            {
              double* dst = dzc + (size_t)(8 * bj + lr) * QP + 2 * lc;
              red_add_f64_keep(dst, ta[0][0] + ta[1][0], keep); red_add_f64_keep(dst + 1, ta[0][1] + ta[1][1], keep);
            }
            if constexpr (QR > 0) {
              if ((c_p & 7) == 0) {
                double* dst = dzr + (size_t)(8 * bi + 2 * half + (c_p >> 3)) * QP + 8 + c_rh * QK;
#pragma unroll
                for (int e = 0; e < QK; ++e) if (8 + c_rh * QK + e < QP) red_add_f64_keep(dst + e, rc[e], keep);
              }
              if (c_p < 8) {
                double* dst = dzc + (size_t)(8 * bj + c_p) * QP + 8 + c_rh * QK;
#pragma unroll
                for (int e = 0; e < QK; ++e) if (8 + c_rh * QK + e < QP) red_add_f64_keep(dst + e, cc[e], keep);
              }
            }
          }
          __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int rr = 0; rr < R; ++rr) drT[(size_t)(8 * bj + k) * RS + lane + 32 * rr] += cs[k][rr];
      }
      __syncthreads();
      wc[0] = wn[0]; wc[1] = wn[1];
    }
    // ---- drain: dv summed over the warps in fixed order, dr transposed back to [row][Mp]
#pragma unroll
    for (int mt = 0; mt < 8; ++mt) {
      xdv[((size_t)warp * QP + 2 * lc) * RS + 8 * mt + lr] = cdv[mt][0];
      xdv[((size_t)warp * QP + 2 * lc + 1) * RS + 8 * mt + lr] = cdv[mt][1];
    }
#pragma unroll
    for (int q = 0; q < QR; ++q)
#pragma unroll
      for (int rr = 0; rr < R; ++rr) xdv[((size_t)warp * QP + 8 + q) * RS + lane + 32 * rr] = dvr[rr][q];
    __syncthreads();
    for (int idx = tid; idx < nc * QP; idx += T) {
      const int row = idx / QP, q = idx - row * QP;
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < kFusedWarps; ++w) a += xdv[((size_t)w * QP + q) * RS + row];
      __stcs(p.dv + ((int64_t)b * p.n + n0) * QP + idx, a);
    }
    {
      double* dst = p.dr + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < nc * p.mp; idx += T) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        __stcs(dst + idx, drT[(size_t)m * RS + row]);
      }
    }
  }
}

}  // namespace dpgp
