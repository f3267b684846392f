// Fused psi2 backward, tensor-core formulation (bwd_variant 3).  Same decomposition, schedule, shared-memory layout,
// second phase and outputs as psi2_bwd_fused_kernel (psi2_bwd_fused.cuh); what changes is the FIRST phase.
//
// psi2_bwd_fused_kernel maps a lane to two rows, so the exponent sum_q v_nq D_pq and the accumulation
// dv_nq += g D_pq are 10 + 10 dependent DFMAs per unit issued from 255 registers by two warps per scheduler:
// the profile is latency-bound (`wait` 2.1 warps per issue, FP64 pipe 49 %, profiles/r01_fused_v3.md).
// Both are small dense contractions over the 8 x 8 (rows x pairs) tiles of a block row:
//     E  [8 rows x 8 pairs]  = (r_m + r_m') + V [8 x Q] . D^T [Q x 8]            3 DMMA m8n8k4 (Q = 10 padded to 12)
//     dV [8 rows x (Q + 1)] += G [8 x 8 pairs] . [D | 1] [8 x (Q + 1)]           4 DMMA (Q + 1 = 11 padded to 16)
// so the first phase runs them on the FP64 tensor cores.  The accumulator layout of the first product (row = lane/4,
// columns 2 (lane%4), 2 (lane%4) + 1) IS the A-operand layout of the second one once the 8 pairs are enumerated as
// column 2t <-> pair t, column 2t+1 <-> pair t + 4: no data movement between exp and the second contraction.
// The "ones" column of [D | 1] delivers the row sums of g (the d r of the block row's m) for free in the padding;
// the column sums stay lane-local because a lane keeps its (row, pair) across the block rows of a block.
// Per unit: 12 (exponent) + 9 (exp) + 16 (dv + row sums) + 1 (column sums) + 10 (dD, second phase) = 48 FP64-pipe
// issue slots against 45, but 3.5 tensor instructions replace 21 scalar ones, every lane carries 2 x TU independent
// exp chains, and v lives in 24 fragment registers instead of 2 x Q.
#pragma once
#include "../psi2_bwd_fused.cuh"

namespace dpgp {

__device__ __forceinline__ void dmma884_f(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Shared-memory layouts of the tensor-core variant.  A fragment access touches (row = 8 t + lane/4, 4 consecutive m or
// pairs = lane%4); the padded [m][65] layout of psi2_bwd_fused_kernel makes those 4-way bank conflicts (17 excess
// wavefronts per unit in the first capture, profiles/r01_fused_tc_v1.md).  Rotating every 64-double row by
// 4 (m & 3) + ((m >> 2) & 3) makes the fragment pattern, the transposing fill / drain (32 consecutive m of one row)
// and the second phase (8 pairs of one parity, one row) all conflict-free.
__device__ __forceinline__ int tc_rot(int m) { return 4 * (m & 3) + ((m >> 2) & 3); }
__device__ __forceinline__ int tc_idx(int m, int row) { return m * 64 + ((row + tc_rot(m)) & 63); }

template <int QP, int EXPV>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) psi2_bwd_tc_kernel(Psi2BwdFusedParams p) {
  extern __shared__ __align__(16) double sm[];
  constexpr int R = 2, RS = 32 * R + 1, ROWS = 32 * R, DS = QP + 2, T = kFusedWarps * 32, PB = kFusedPB;
  constexpr int QH = QP / 2, QHP = (QH + 1) & ~1;
  constexpr int KS = (QP + 3) / 4;                      // k-steps of the exponent product
  constexpr int NT = (QP + 1 + 7) / 8;                  // column tiles of [D | 1]
  constexpr int NRT = ROWS / 8;                         // row tiles
  constexpr int TU = 2;                                 // row tiles in flight
  constexpr int ONE_J = QP / 8, ONE_T = (QP % 8) / 2, ONE_E = QP % 2;      // where the ones column lands in the fragments
  static_assert(NT <= 2 && KS <= 3, "tensor-core variant is instantiated for QP <= 12");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lc = lane & 3;
  double* rT = sm;
  double* drT = rT + (size_t)p.mp * RS;
  double* zs = drT + (size_t)p.mp * RS;
  double* etab = zs + (size_t)p.mp * QP;
  double* vt = etab + kExpTabSize;                      // [ROWS][2][QHP]
  double* dtab = vt + (size_t)ROWS * 2 * QHP;
  double* gtab = dtab + (size_t)kFusedWarps * PB * DS;
  double* xdv = gtab;
  double* dtw = dtab + (size_t)warp * PB * DS;
  double* gtw = gtab + (size_t)warp * PB * RS;          // [16 pairs][64], rows rotated (tc_idx)

  for (int i = tid; i < p.mp * QP; i += T) { const int m = i / QP, q = i % QP; zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0; }
  load_exp_table(etab, p.exptab);
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  Exp<EXPV> ex; ex.init(etab);
  const uint64_t keep = l2_evict_last_policy();
  const size_t slice_len = (size_t)p.nrounds * kFusedWarps * 64 * QP;
  const int p2_pair = lane >> 1, p2_qh = lane & 1;
  const int p2_pp = (lane >> 1) & 7, p2_rh = lane >> 4;
  const int bpair = (lr >> 1) + 4 * (lr & 1);           // pair enumerated by B-operand column lane/4: col 2t <-> t, 2t+1 <-> t+4

  const int64_t items = p.ngroups * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  int cur_b = -1, seg = -1;
  double* mypart = nullptr;
  for (int64_t item = lo; item < hi; ++item) {
    const int b = (int)(item / p.ngroups);
    const int64_t n0 = (item % p.ngroups) * ROWS;
    const int nc = (int)min((int64_t)ROWS, p.n - n0);
    if (b != cur_b) {
      cur_b = b; ++seg;
      mypart = p.part + ((size_t)blockIdx.x * p.nseg + seg) * slice_len;
      if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = b;
    }
    __syncthreads();
    {
      const double* src = p.r + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < ROWS * p.mp; idx += T) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        rT[tc_idx(m, row)] = (row < nc) ? __ldcs(src + idx) : kRClamp;
      }
      for (int idx = tid; idx < p.mp * 64; idx += T) drT[idx] = 0.0;
      const double* vsrc = p.v + ((int64_t)b * p.n + n0) * QP;
      for (int idx = tid; idx < ROWS * 2 * QHP; idx += T) {
        const int row = idx / (2 * QHP), rem = idx - row * 2 * QHP, h = rem / QHP, j = rem - h * QHP;
        vt[idx] = (row < nc && j < QH) ? __ldcs(vsrc + row * QP + h * QH + j) : 0.0;
      }
    }
    // A-operand fragments of V (row = 8 t + lane/4, q = lane%4 + 4 s) and the dv accumulators of this row group
    double va[NRT][KS], dvc[NRT][NT][2];
#pragma unroll
    for (int t = 0; t < NRT; ++t) {
      const int row = 8 * t + lr;
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const int q = lc + 4 * s;
        va[t][s] = (row < nc && q < QP) ? __ldcs(p.v + ((int64_t)b * p.n + n0 + row) * QP + (q < QP ? q : 0)) : 0.0;
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) { dvc[t][j][0] = 0.0; dvc[t][j][1] = 0.0; }
    }
    const double* gb = p.gbar + (size_t)b * p.m * p.m;
    auto load_w = [&](unsigned short it, double (&w)[2]) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int id = lane + 32 * e;
        w[e] = (it == kSchedIdle) ? 0.0 : sym_cotangent(gb, 8 * (it >> 8) + (id >> 3), 8 * (it & 255) + (id & 7), p.m);
      }
    };
    double wc[2], wn[2] = {0.0, 0.0};
    load_w(p.sched[warp], wc);
    __syncthreads();

    for (int round = 0; round < p.nrounds; ++round) {
      const unsigned short it = p.sched[round * kFusedWarps + warp];
      if (round + 1 < p.nrounds) load_w(p.sched[(round + 1) * kFusedWarps + warp], wn);
      if (it != kSchedIdle) {
        const int bi = it >> 8, bj = it & 255;
        double* slot = mypart + ((size_t)(round * kFusedWarps + warp) * 64) * QP;
        // column sums of g: the lane keeps (row 8t + lane/4, pairs lane%4 and lane%4 + 4) over the 8 block rows
        double cs[NRT][2];
#pragma unroll
        for (int t = 0; t < NRT; ++t) { cs[t][0] = 0.0; cs[t][1] = 0.0; }
        const int mc0 = 8 * bj + lc, mc1 = mc0 + 4;                        // m' of the lane's two pairs
#pragma unroll 1
        for (int half = 0; half < 64 / PB; ++half) {
          {
            const int i = 2 * half + (p2_pair >> 3), k = p2_pair & 7, m = 8 * bi + i, c = 8 * bj + k;
#pragma unroll
            for (int j = 0; j < QH; ++j) {
              const int q = p2_qh * QH + j;
              const double d = zs[m * QP + q] - zs[c * QP + q];
              dtw[p2_pair * DS + q] = d * d;
            }
            const double wv = __shfl_sync(0xffffffffu, (half & 2) ? wc[1] : wc[0], 16 * (half & 1) + p2_pair);
            if (p2_qh == 0) { dtw[p2_pair * DS + QP] = wv; dtw[p2_pair * DS + QP + 1] = 0.0; }
          }
          __syncwarp();
          // ---- phase 1 on the tensor cores: one block row (8 pairs) x 64 rows at a time
#pragma unroll 1
          for (int i2 = 0; i2 < 2; ++i2) {
            const int i = 2 * half + i2;
            const double* dt = dtw + (size_t)(i2 * 8) * DS;
            // B operands: exponent  D[pair(lane/4)][q = lane%4 + 4 s];  second product  [D | 1][pair = lane%4 + 4 s][col = lane/4 + 8 j]
            double bfe[KS], bfd[2][NT], w2[2];
#pragma unroll
            for (int s = 0; s < KS; ++s) { const int q = lc + 4 * s; bfe[s] = (q < QP) ? dt[bpair * DS + q] : 0.0; }
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
              for (int j = 0; j < NT; ++j) {
                const int col = lr + 8 * j;
                bfd[s][j] = (col < QP) ? dt[(lc + 4 * s) * DS + col] : (col == QP ? 1.0 : 0.0);
              }
            w2[0] = dt[lc * DS + QP]; w2[1] = dt[(lc + 4) * DS + QP];
            const int mi = 8 * bi + i, gp0i = i2 * 8 + lc, gp1i = gp0i + 4;
#pragma unroll
            for (int t0 = 0; t0 < NRT; t0 += TU) {
              double e[2 * TU], w[2 * TU], g[2 * TU];
#pragma unroll
              for (int u = 0; u < TU; ++u) {
                const int t = t0 + u;
                const double rm = rT[tc_idx(mi, 8 * t + lr)];
                double c[2] = {rm + rT[tc_idx(mc0, 8 * t + lr)], rm + rT[tc_idx(mc1, 8 * t + lr)]};
#pragma unroll
                for (int s = 0; s < KS; ++s) dmma884_f(c, va[t][s], bfe[s]);
                e[2 * u] = c[0]; e[2 * u + 1] = c[1]; w[2 * u] = w2[0]; w[2 * u + 1] = w2[1];
              }
              exp_scaled_k<EXPV, 2 * TU>(ex, e, w, g);
#pragma unroll
              for (int u = 0; u < TU; ++u) {
                const int t = t0 + u;
                gtw[tc_idx(gp0i, 8 * t + lr)] = g[2 * u]; gtw[tc_idx(gp1i, 8 * t + lr)] = g[2 * u + 1];
                cs[t][0] += g[2 * u]; cs[t][1] += g[2 * u + 1];
#pragma unroll
                for (int j = 0; j < NT; ++j) { dmma884_f(dvc[t][j], g[2 * u], bfd[0][j]); dmma884_f(dvc[t][j], g[2 * u + 1], bfd[1][j]); }
              }
            }
            // row sums of g of this block row sit in the ones column of the accumulators: move them to d r and clear
            if (lc == ONE_T) {
#pragma unroll
              for (int t = 0; t < NRT; ++t) {
                drT[tc_idx(8 * bi + i, 8 * t + lr)] += dvc[t][ONE_J][ONE_E];
                dvc[t][ONE_J][ONE_E] = 0.0;
              }
            }
          }
          __syncwarp();
          // ---- phase 2 (unchanged): lane <-> (two pairs, q half, half of the rows)
          {
            constexpr int HR = ROWS / 2;
            double acc0[QH], acc1[QH];
#pragma unroll
            for (int j = 0; j < QH; ++j) { acc0[j] = 0.0; acc1[j] = 0.0; }
            const double* gp0 = gtw + (size_t)(2 * p2_pp) * 64;
            const double* gp1 = gp0 + 64;
            const int rot0 = tc_rot(2 * p2_pp) + p2_rh * HR, rot1 = tc_rot(2 * p2_pp + 1) + p2_rh * HR;
            const double* vp = vt + (size_t)(p2_rh * HR) * 2 * QHP + p2_qh * QHP;
#pragma unroll 4
            for (int rw = 0; rw < HR; ++rw) {
              const int row = (rw + p2_rh) & (HR - 1);
              const double g0 = gp0[(row + rot0) & 63], g1 = gp1[(row + rot1) & 63];
              double vv[QHP];
#pragma unroll
              for (int j = 0; j < QHP; j += 2) { const double2 t2 = *reinterpret_cast<const double2*>(vp + (size_t)row * 2 * QHP + j); vv[j] = t2.x; vv[j + 1] = t2.y; }
#pragma unroll
              for (int j = 0; j < QH; ++j) { acc0[j] = fma(g0, vv[j], acc0[j]); acc1[j] = fma(g1, vv[j], acc1[j]); }
            }
#pragma unroll
            for (int j = 0; j < QH; ++j) {
              acc0[j] += __shfl_down_sync(0xffffffffu, acc0[j], 16);
              acc1[j] += __shfl_down_sync(0xffffffffu, acc1[j], 16);
            }
            if (p2_rh == 0) {
              double* dst = slot + (size_t)(half * PB + 2 * p2_pp) * QP + p2_qh * QH;
#pragma unroll
              for (int j = 0; j < QH; ++j) { red_add_f64_keep(dst + j, acc0[j], keep); red_add_f64_keep(dst + QP + j, acc1[j], keep); }
            }
          }
          __syncwarp();
        }
#pragma unroll
        for (int t = 0; t < NRT; ++t) {
          drT[tc_idx(8 * bj + lc, 8 * t + lr)] += cs[t][0];
          drT[tc_idx(8 * bj + lc + 4, 8 * t + lr)] += cs[t][1];
        }
      }
      __syncthreads();
      wc[0] = wn[0]; wc[1] = wn[1];
    }
    // ---- drain: accumulator fragments -> xdv[warp][q][row], summed over the warps in fixed order; d r transposed back
#pragma unroll
    for (int t = 0; t < NRT; ++t)
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int q = 8 * j + 2 * lc + e;
          if (q < QP) xdv[((size_t)warp * QP + q) * RS + 8 * t + lr] = dvc[t][j][e];
        }
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
        __stcs(dst + idx, drT[tc_idx(m, row)]);
      }
    }
  }
}

}  // namespace dpgp
