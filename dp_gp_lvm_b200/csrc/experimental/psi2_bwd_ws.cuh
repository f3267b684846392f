// Fused psi2 backward, warp-specialised (bwd_variant 5).  Same mathematics, schedule, slices and outputs as
// psi2_bwd_fused_kernel (psi2_bwd_fused.cuh); the two phases of that kernel run on DIFFERENT warps here:
//
//   warps 0..7   producers: phase 1 (exponent, exp, dv, row / column sums of g) for their block of the round, one block row
//                (8 pairs x 64 rows) at a time into a double-buffered g tile, ~200 registers (setmaxnreg.inc)
//   warps 8..15  helpers: helper w consumes producer w's tiles -- phase 2, dD_pq += sum_rows g[p][row] v[row][q] -- and adds
//                the totals into the CTA's dD slice, 56 registers (setmaxnreg.dec)
//
// Why: the single-role kernel is latency-bound with two warps per scheduler at 255 registers (FP64 pipe 52 %), and every
// variant that bought occupancy by shrinking the row group paid for it in shared-memory traffic (D is re-read per row).
// The second phase is an FMA-dense, register-light stream; giving it its own warps raises the warps per scheduler from 2
// to 4 WITHOUT changing the per-unit shared-memory traffic of phase 1.  Hand-over: mbarrier pairs full / empty per
// (producer, buffer); every lane arrives, so no separate fence is needed.
#pragma once
#include "../psi2_bwd_fused.cuh"

namespace dpgp {

constexpr int kWsProducerRegs = 200, kWsHelperRegs = 56;

template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// smem (doubles): rT[mp*RS] | drT[mp*RS] | zs[mp*QP] | etab[256] | vt[64][2*QHP] | dtab[8][16*(QP+2)] | gt[8][2][8*RS]
//                 (gt aliases xdv[8][QP][RS] in the drain) | full[8][2], empty[8][2] mbarriers
template <int QP>
__host__ __device__ inline size_t ws_smem_bytes(int mp) { return fused_smem_bytes<QP, 2, 1>(mp) + 4 * kFusedWarps * 8; }

template <int QP, int EXPV, int KU, bool PRODUCER>
__device__ __forceinline__ void ws_role_loop(const Psi2BwdFusedParams& p, double* rT, double* drT, double* zs, double* etab, double* vt,
                                             double* dtw, double* gtw, double* xdv, uint64_t* full, uint64_t* empty, int wt) {
  constexpr int R = 2, RS = 32 * R + 1, ROWS = 32 * R, DS = QP + 2, T = kFusedWarps * 64, PB = kFusedPB;
  constexpr int QH = QP / 2, QHP = (QH + 1) & ~1;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t keep = l2_evict_last_policy();
  const size_t slice_len = (size_t)p.nrounds * kFusedWarps * 64 * QP;
  const int64_t items = p.ngroups * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  int cur_b = -1, seg = -1;
  unsigned tcnt = 0;                                    // hand-over tiles so far: same sequence on both sides
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
    __syncthreads();                                    // (1) previous drain finished with rT / drT / xdv / vt
    {
      const double* src = p.r + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < ROWS * p.mp; idx += T) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        rT[(size_t)m * RS + row] = (row < nc) ? __ldcs(src + idx) : kRClamp;
      }
      for (int idx = tid; idx < p.mp * RS; idx += T) drT[idx] = 0.0;
      const double* vsrc = p.v + ((int64_t)b * p.n + n0) * QP;
      for (int idx = tid; idx < ROWS * 2 * QHP; idx += T) {
        const int row = idx / (2 * QHP), rem = idx - row * 2 * QHP, h = rem / QHP, j = rem - h * QHP;
        vt[idx] = (row < nc && j < QH) ? __ldcs(vsrc + row * QP + h * QH + j) : 0.0;
      }
    }
    __syncthreads();                                    // (2) tiles staged

    if constexpr (PRODUCER) {
      // ================================================================================ producers: phase 1
      Exp<EXPV> ex; ex.init(etab);
      const int p2_pair = lane >> 1, p2_qh = lane & 1;
      double vq[R][QP], dv[R][QP];
#pragma unroll
      for (int rr = 0; rr < R; ++rr) {
        const int row = lane + 32 * rr;
#pragma unroll
        for (int q = 0; q < QP; ++q) {
          // v of the lane's rows from the staged tile: vt[row][half][QHP]
          vq[rr][q] = vt[(size_t)row * 2 * QHP + (q / QH) * QHP + (q % QH)];
          dv[rr][q] = 0.0;
        }
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
      load_w(p.sched[wt], wc);
      for (int round = 0; round < p.nrounds; ++round) {
        const unsigned short it = p.sched[round * kFusedWarps + wt];
        if (round + 1 < p.nrounds) load_w(p.sched[(round + 1) * kFusedWarps + wt], wn);
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
#pragma unroll 1
            for (int i2 = 0; i2 < 2; ++i2, ++tcnt) {
              const int i = 2 * half + i2, buf = tcnt & 1;
              if (tcnt >= 2) mbar_wait(&empty[buf], ((tcnt >> 1) - 1) & 1);       // the helper has drained this buffer
              double* gbuf = gtw + (size_t)buf * 8 * RS + lane;
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
#pragma unroll
                  for (int rr = 0; rr < R; ++rr) gbuf[(size_t)(k0 + u) * RS + 32 * rr] = g[u * R + rr];
#pragma unroll
                  for (int q = 0; q < QP; ++q)
#pragma unroll
                    for (int rr = 0; rr < R; ++rr) dv[rr][q] = fma(g[u * R + rr], dq[u][q], dv[rr][q]);
#pragma unroll
                  for (int rr = 0; rr < R; ++rr) { rs[rr] += g[u * R + rr]; cs[k0 + u][rr] += g[u * R + rr]; }
                }
              }
#pragma unroll
              for (int rr = 0; rr < R; ++rr) drT[(size_t)(8 * bi + i) * RS + lane + 32 * rr] += rs[rr];
              mbar_arrive(&full[buf]);                  // every lane: its stores to the tile are published
            }
            __syncwarp();                               // the pair table is rebuilt next
          }
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int rr = 0; rr < R; ++rr) drT[(size_t)(8 * bj + k) * RS + lane + 32 * rr] += cs[k][rr];
        }
        team_barrier(0, kFusedWarps * 32);              // producers only: the round's d r updates are in place
        wc[0] = wn[0]; wc[1] = wn[1];
      }
      __syncthreads();                                  // (3) helpers are done with the g tiles
#pragma unroll
      for (int q = 0; q < QP; ++q)
#pragma unroll
        for (int rr = 0; rr < R; ++rr) xdv[((size_t)wt * QP + q) * RS + lane + 32 * rr] = dv[rr][q];
    } else {
      // ================================================================================ helpers: phase 2
      const int h_rh = lane >> 4, h_p8 = (lane >> 1) & 7, h_qh = lane & 1;
      constexpr int HR = ROWS / 2;
      const double* vp = vt + (size_t)(h_rh * HR) * 2 * QHP + h_qh * QHP;
      for (int round = 0; round < p.nrounds; ++round) {
        const unsigned short it = p.sched[round * kFusedWarps + wt];
        if (it == kSchedIdle) continue;
        double* slot = mypart + ((size_t)(round * kFusedWarps + wt) * 64) * QP + (size_t)h_p8 * QP + h_qh * QH;
#pragma unroll 1
        for (int t8 = 0; t8 < 8; ++t8, ++tcnt) {
          const int buf = tcnt & 1;
          mbar_wait(&full[buf], (tcnt >> 1) & 1);
          const double* gp = gtw + (size_t)buf * 8 * RS + (size_t)h_p8 * RS + h_rh * HR;
          double acc[QH];
#pragma unroll
          for (int j = 0; j < QH; ++j) acc[j] = 0.0;
#pragma unroll 4
          for (int rw = 0; rw < HR; ++rw) {
            const int row = (rw + h_rh) & (HR - 1);
            const double gl = gp[row];
            double vv[QHP];
#pragma unroll
            for (int j = 0; j < QHP; j += 2) { const double2 t2 = *reinterpret_cast<const double2*>(vp + (size_t)row * 2 * QHP + j); vv[j] = t2.x; vv[j + 1] = t2.y; }
#pragma unroll
            for (int j = 0; j < QH; ++j) acc[j] = fma(gl, vv[j], acc[j]);
          }
          mbar_arrive(&empty[buf]);                     // every lane: its reads of the tile are complete (consumed above)
#pragma unroll
          for (int j = 0; j < QH; ++j) acc[j] += __shfl_down_sync(0xffffffffu, acc[j], 16);
          if (h_rh == 0) {
            double* dst = slot + (size_t)t8 * 8 * QP;
#pragma unroll
            for (int j = 0; j < QH; ++j) red_add_f64_keep(dst + j, acc[j], keep);
          }
        }
      }
      __syncthreads();                                  // (3)
    }
    __syncthreads();                                    // (4) dv partials of the 8 producers are in xdv
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

template <int QP, int EXPV, int KU = DPGP_FUSED_KU>
__global__ void __launch_bounds__(kFusedWarps * 64, 1) psi2_bwd_ws_kernel(Psi2BwdFusedParams p) {
  extern __shared__ __align__(16) double sm[];
  constexpr int R = 2, RS = 32 * R + 1, ROWS = 32 * R, DS = QP + 2, T = kFusedWarps * 64, PB = kFusedPB;
  constexpr int QH = QP / 2, QHP = (QH + 1) & ~1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp < kFusedWarps;
  const int wt = warp % kFusedWarps;                     // slot in the round schedule (producer) / producer served (helper)
  double* rT = sm;
  double* drT = rT + (size_t)p.mp * RS;
  double* zs = drT + (size_t)p.mp * RS;
  double* etab = zs + (size_t)p.mp * QP;
  double* vt = etab + kExpTabSize;                      // [ROWS][2][QHP]
  double* dtab = vt + (size_t)ROWS * 2 * QHP;
  double* gtab = dtab + (size_t)kFusedWarps * PB * DS;
  double* xdv = gtab;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gtab + (size_t)kFusedWarps * PB * RS);
  uint64_t* full = bars + wt * 2;                       // [2] per producer
  uint64_t* empty = bars + 2 * kFusedWarps + wt * 2;
  double* dtw = dtab + (size_t)wt * PB * DS;
  double* gtw = gtab + (size_t)wt * PB * RS;            // two buffers of 8 pairs x RS

  for (int i = tid; i < p.mp * QP; i += T) { const int m = i / QP, q = i % QP; zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0; }
  load_exp_table(etab, p.exptab);
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  if (tid == 0) {
    for (int i = 0; i < 4 * kFusedWarps; ++i) mbar_init(&bars[i], 32);
    mbar_fence_init();
  }
  __syncthreads();
  // The two roles run completely separate copies of the item loop: ptxas budgets registers per region only when the
  // code after setmaxnreg is not shared between the roles.
  if (producer) {
    setmaxnreg_inc<kWsProducerRegs>();
    ws_role_loop<QP, EXPV, KU, true>(p, rT, drT, zs, etab, vt, dtw, gtw, xdv, full, empty, wt);
  } else {
    setmaxnreg_dec<kWsHelperRegs>();
    ws_role_loop<QP, EXPV, KU, false>(p, rT, drT, zs, etab, vt, dtw, gtw, xdv, full, empty, wt);
  }
}

}  // namespace dpgp
