// Fused backward of the psi2 statistic: ONE exp per (cluster, row, pair m <= m') unit feeds all three
// reductions (replaces TensorFlow autodiff through src/kernels/rbf_kernel.py:164-199; supersedes the two
// kernels of psi2_bwd.cuh, which evaluate the exponential twice).
//
//   g_np  = Gs_p exp(r_nm + r_nm' + sum_q v_nq D_pq)           (psi2.cuh; Gs = symmetrised cotangent of Psi2)
//   dr_nm = sum_{m'} g_n(m,m')     dv_nq = sum_p g_np D_pq      dD_pq = sum_n g_np v_nq
//
// Work decomposition.  All 8 warps of a CTA work on the same group of 32 R rows of one cluster.  The M/8 x M/8
// triangle of 8x8 pair blocks is walked in ROUNDS built on the host from a round-robin tournament
// (1-factorisation of the complete graph on the m-blocks): the <= 8 blocks of a round touch disjoint m-blocks
// and go to one warp each.  Hence dr accumulates in ONE shared [Mp][rows] array without conflicts (a CTA
// barrier separates rounds) and in a fixed order: results are bitwise reproducible.
//
// A warp handles its block 16 pairs (two block rows) at a time, in two phases:
//   phase 1  lane <-> R rows.  exponent, exp, dv (registers, kept over all rounds), row / column sums of g
//            (dr); g itself goes to a warp-private shared tile gt[16 pairs][rows].
//   phase 2  lane <-> (two pairs, half of the q range, half of the rows).  dD_pq += sum_rows gt[p][row] v[row][q]
//            with the accumulators in registers (a 2 x Q/2 register tile: 7 shared-memory wavefronts per 10 FMAs);
//            the two row halves are combined by one shuffle per accumulator per 16 pairs.  There is no
//            per-unit cross-lane reduction anywhere (a shuffle-based transposing reduction cost 45 non-FP64
//            issues and 6 FP64 issues per unit in the first version, profiles/r01_psi2.md).
// The dD totals of a 16-pair step leave the kernel through red.global.add.f64 into global slices; every slice address
// is only ever updated by one lane of one warp, in program order -> deterministic.  Two layouts (template flag DZ):
// per-CTA dD slices summed by dd_fused_reduce_kernel (bwd_variant 1), or -- the default, bwd_variant 6 -- dD contracted
// with 2 (z_m - z_m') on the spot into per-warp dz slices summed by dz_fused_reduce_kernel (see the DZ comment below).
// FP64-pipe issues per unit at Q = 10: 11 (exponent) + 9 (table exp incl. weight) + 10 (dv) + 2 (dr) +
// 10 (dD) + 0.3 (pair table) = 42 (measured 45: diagonal blocks are swept in full, branch-free), against
// 2 x 26 + 22 = 74 for the two-kernel version.  Measured at 262 144 rows, Q = 10, M = 128, 10 clusters: 96.8 ms with dD
// slices, 98.0 ms with dz folding (FP64 pipe 52 %, shared-memory pipe 66 % in profiles/r01_fused_v4_ku2.md, taken at
// 100.9 ms before the phase-2 loop was fully unrolled); variants that traded one of the two for the other (tensor-core
// first phase, 16 warps per SM, warp specialisation) are kept selectable and documented in profiles/r01_psi2.md.
#pragma once
#include "common.cuh"

namespace dpgp {

// Symmetrised cotangent of Psi2 for the pair (m <= c): Gbar[m,c] + Gbar[c,m] off the diagonal, Gbar[m,m] on it; 0 for
// pairs below the diagonal or outside M (padding), so that dead pair slots contribute g = 0.
__device__ __forceinline__ double sym_cotangent(const double* gb, int m, int c, int M) {
  if (m >= M || c >= M || m > c) return 0.0;
  if (m == c) return gb[(size_t)m * M + m];
  return gb[(size_t)m * M + c] + gb[(size_t)c * M + m];
}

constexpr int kFusedWarps = 8;
constexpr int kFusedPB = 16;                    // pairs per phase-1 / phase-2 hand-over (two block rows)
constexpr unsigned short kSchedIdle = 0xffff;
#ifndef DPGP_SLICE_RMW
#define DPGP_SLICE_RMW 0
#endif
// dD slice update: 0 = red.global.add.f64 (default), 1 = ld.cg / st.cg read-modify-write.  Measured at 262 144 rows:
// RED 108 ms and 28.7 GB of DRAM traffic, RMW 127 ms and 40 GB -- the 103 MB of slices do not stay in L2 either way.
constexpr bool kSliceRmw = DPGP_SLICE_RMW != 0;
// Unroll factor of the phase-2 row loop (32 rows per lane at R = 2).  Measured at 262 144 rows, slices / dz variant:
// 1: 111.4, 2: 104.3, 4: 100.0, 8: 98.0, 16: 97.5 / 101.0, 32 (full): 96.8 / 98.0 ms.
#ifndef DPGP_XP_P2_UNROLL
#define DPGP_XP_P2_UNROLL 32
#endif
// Unroll of the phase-1 block-row loop: 2 spills (255 registers) and runs 123.8 ms; 1 (default) 97.9 ms.
#ifndef DPGP_XP_I2_UNROLL
#define DPGP_XP_I2_UNROLL 1
#endif
#define DPGP_PRAGMA_(x) _Pragma(#x)
#define DPGP_UNROLL(n) DPGP_PRAGMA_(unroll n)

struct Psi2BwdFusedParams {
  const double* r; const double* v; const double* z; const double* gbar; const double* exptab;
  const unsigned short* sched;   // [nrounds][8]: (bi << 8) | bj, or kSchedIdle
  double* dr;                    // [B,N,Mp]   may alias r (a CTA stages its rows before it overwrites them)
  double* dv;                    // [B,N,QP]
  double* part;                  // [grid*nseg][nrounds*8*64*QP]   zero on entry
  int* tags;                     // [grid*nseg]
  int64_t n; int q, m, mp, b, nrounds, nseg; int64_t ngroups;
};

__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
// Same with an L2 evict-last policy: the per-CTA dD slices (0.7 MB each) are re-visited once per row group and
// should stay L2-resident while the r / v stream (evict-first loads) passes through.
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void red_add_f64_keep(double* addr, double v, uint64_t pol) {
  asm volatile("red.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"(v), "l"(pol) : "memory");
}

// smem (doubles): rT[mp*RS] | drT[mp*RS] | zs[mp*QP] | etab[256] | vt[ROWS][2*QHP] | dtab[8][16*(QP+2)] |
//                 gt[8][16*RS]  (gt aliases xdv[8][QP][RS] during the drain)
// TEAMS = 2: two teams of 8 warps share a 32-row group (R = 1) and split the rounds; every team has its own d r
// accumulator (summed in the drain), every warp its own pair table and g tile.
template <int QP, int R, int TEAMS = 1>
__host__ __device__ inline size_t fused_smem_bytes(int mp) {
  const int RS = 32 * R + 1, ROWS = 32 * R, QHP = (QP / 2 + 1) & ~1, NW = kFusedWarps * TEAMS;
  const size_t gt = (size_t)NW * kFusedPB * RS, xd = (size_t)NW * QP * RS;
  return ((1 + TEAMS) * (size_t)mp * RS + (size_t)mp * QP + kExpTabSize + (size_t)ROWS * 2 * QHP +
          (size_t)NW * kFusedPB * (QP + 2) + (gt > xd ? gt : xd)) * 8;
}
__device__ __forceinline__ void team_barrier(int team, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(threads) : "memory");
}

// Pair steps advanced in lockstep in the first phase.  Measured at 262 144 rows: KU = 1 107.6 ms, 2 100.8 ms, 4 101.0 ms,
// 8 101.5 ms (ptxas serialises the exp chains of consecutive pair steps unless the source interleaves them).
#ifndef DPGP_FUSED_KU
#define DPGP_FUSED_KU 2
#endif
// DZ = true (bwd_variant 6, the default): the dD totals of a 16-pair step are contracted with 2 (z_m - z_m') on the spot,
// dz_m += 2 d dD, dz_m' -= 2 d dD (the only consumer of dD, bound.cuh: zchain_kernel), and added into two [Mp][QP] slices
// per WARP (row side, column side) instead of a [rounds][8][64][QP] slice per CTA: 20 KB per warp, 24 MB in all, which
// stays in L2 (the 103 MB of dD slices did not: 142.7 GB of DRAM traffic per launch at N = 1M against 22 GB of inputs
// and outputs).  Every slice address has one writer lane, in program order -> still bitwise reproducible.
// Price: ~157 more warp instructions per 16-pair step (the (z_m - z_m') factors, 52 shuffles of the two cross-lane
// sums, index arithmetic) = +6.3 % instructions: 98.0 ms against 96.8 ms for the slices (bwd_variant 1) at 262 144 rows;
// DRAM traffic per launch at 65 536 rows 1.41 GB against 8.9 GB (profiles/r01_fused_dz.md).  1 % of kernel time buys the
// algorithmic traffic, ~180 MB less workspace and two launches less per evaluation, so this is the default.
// The dz-folding variant is compiled with four pair steps in lockstep (KU = 1 / 2 / 4 / 8: 107.1 / 103.7 / 98.0 / 99.0 ms with
// the phase-2 loop fully unrolled; 113.0 / 108.6 / 103.8 / 105.1 ms at unroll 4).
#ifndef DPGP_FUSED_KU_DZ
#define DPGP_FUSED_KU_DZ 4
#endif
constexpr int kFusedKuDz = DPGP_FUSED_KU_DZ;
template <int QP, int EXPV, int R, int TEAMS = 1, int KU = DPGP_FUSED_KU, bool DZ = false>
__global__ void __launch_bounds__(kFusedWarps * 32 * TEAMS, 1) psi2_bwd_fused_kernel(Psi2BwdFusedParams p) {
  static_assert(!DZ || TEAMS == 1, "dz folding: one team");
  extern __shared__ __align__(16) double sm[];
  constexpr int RS = 32 * R + 1, ROWS = 32 * R, DS = QP + 2, T = kFusedWarps * 32 * TEAMS, PB = kFusedPB;
  constexpr int NW = kFusedWarps * TEAMS;
  static_assert(TEAMS == 1 || R == 1, "two teams: 32-row groups");
  constexpr int QH = QP / 2, QHP = (QH + 1) & ~1;       // q range split in two halves for phase 2
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int team = warp / kFusedWarps, wt = warp % kFusedWarps;       // team and slot of the warp in the round schedule
  double* rT = sm;
  double* drT0 = rT + (size_t)p.mp * RS;                // [TEAMS][mp][RS]
  double* drT = drT0 + (size_t)team * p.mp * RS;
  double* zs = drT0 + (size_t)TEAMS * p.mp * RS;
  double* etab = zs + (size_t)p.mp * QP;
  double* vt = etab + kExpTabSize;                      // [ROWS][2][QHP]
  double* dtab = vt + (size_t)ROWS * 2 * QHP;
  double* gtab = dtab + (size_t)NW * PB * DS;
  double* xdv = gtab;                                   // alias, used only between the last round and the next fill
  double* dtw = dtab + (size_t)warp * PB * DS;
  double* gtw = gtab + (size_t)warp * PB * RS;

  for (int i = tid; i < p.mp * QP; i += T) { const int m = i / QP, q = i % QP; zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0; }
  load_exp_table(etab, p.exptab);
  if (!DZ) for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  Exp<EXPV> ex; ex.init(etab);
  const uint64_t keep = l2_evict_last_policy();
  double* dzr = DZ ? p.part + ((size_t)blockIdx.x * kFusedWarps + wt) * 2 * p.mp * QP : nullptr;      // row-side slice of this warp
  double* dzc = DZ ? dzr + (size_t)p.mp * QP : nullptr;                                               // column side
  const size_t slice_len = (size_t)p.nrounds * kFusedWarps * 64 * QP;
  const int p2_pair = lane >> 1, p2_qh = lane & 1;      // pair table build: lane <-> (pair of the half, q half)
  const int p2_pp = (lane >> 1) & 7, p2_rh = lane >> 4; // phase 2: lane <-> (two pairs, q half, half of the rows)

  const int64_t items = p.ngroups * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  int cur_b = -1, seg = -1;
  double* mypart = nullptr;
  for (int64_t item = lo; item < hi; ++item) {
    const int b = (int)(item / p.ngroups);
    const int64_t n0 = (item % p.ngroups) * ROWS;
    const int nc = (int)min((int64_t)ROWS, p.n - n0);
    if (!DZ && b != cur_b) {
      cur_b = b; ++seg;
      mypart = p.part + ((size_t)blockIdx.x * p.nseg + seg) * slice_len;
      if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = b;
    }
    __syncthreads();                                    // previous group's drain has finished with rT / drT / xdv
    {
      const double* src = p.r + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < ROWS * p.mp; idx += T) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        rT[(size_t)m * RS + row] = (row < nc) ? __ldcs(src + idx) : kRClamp;      // dead rows: exp(2 kRClamp) is exactly 0 in every variant
      }
      for (int idx = tid; idx < TEAMS * p.mp * RS; idx += T) drT0[idx] = 0.0;
      const double* vsrc = p.v + ((int64_t)b * p.n + n0) * QP;
      for (int idx = tid; idx < ROWS * 2 * QHP; idx += T) {
        const int row = idx / (2 * QHP), rem = idx - row * 2 * QHP, h = rem / QHP, j = rem - h * QHP;
        vt[idx] = (row < nc && j < QH) ? __ldcs(vsrc + row * QP + h * QH + j) : 0.0;
      }
    }
    double vq[R][QP], dv[R][QP];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      const int row = lane + 32 * rr;
      const double* vs = p.v + ((int64_t)b * p.n + n0 + (row < nc ? row : 0)) * QP;
#pragma unroll
      for (int q = 0; q < QP; q += 2) {
        const double2 t2 = __ldcs(reinterpret_cast<const double2*>(vs + q));
        vq[rr][q] = (row < nc) ? t2.x : 0.0; vq[rr][q + 1] = (row < nc) ? t2.y : 0.0;
        dv[rr][q] = 0.0; dv[rr][q + 1] = 0.0;
      }
    }
    const double* gb = p.gbar + (size_t)b * p.m * p.m;
    // symmetrised cotangents of a block's 64 pairs, two per lane (pair ids lane, lane + 32); those of the next
    // round are loaded while the current round computes (the L2 latency showed as 12 % of the samples otherwise)
    auto load_w = [&](unsigned short it, double (&w)[2]) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int id = lane + 32 * e;
        w[e] = (it == kSchedIdle) ? 0.0 : sym_cotangent(gb, 8 * (it >> 8) + (id >> 3), 8 * (it & 255) + (id & 7), p.m);
      }
    };
    double wc[2], wn[2] = {0.0, 0.0};
    load_w(team < p.nrounds ? p.sched[team * kFusedWarps + wt] : kSchedIdle, wc);
    __syncthreads();

    for (int round = team; round < p.nrounds; round += TEAMS) {          // the teams take alternate rounds
      const unsigned short it = p.sched[round * kFusedWarps + wt];
      if (round + TEAMS < p.nrounds) load_w(p.sched[(round + TEAMS) * kFusedWarps + wt], wn);
      if (it != kSchedIdle) {
        const int bi = it >> 8, bj = it & 255;
        double* slot = mypart + ((size_t)(round * kFusedWarps + wt) * 64) * QP;
        const double* rcol = rT + (size_t)(8 * bj) * RS + lane;
        double cs[8][R];
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int rr = 0; rr < R; ++rr) cs[k][rr] = 0.0;
#pragma unroll 1
        for (int half = 0; half < 64 / PB; ++half) {
          // ---- table of this half's 16 pairs: D[pair][q] and the symmetrised cotangent; lane <-> (pair, q half)
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
          // ---- phase 1: lane <-> rows
          DPGP_UNROLL(DPGP_XP_I2_UNROLL)
          for (int i2 = 0; i2 < 2; ++i2) {
            const int i = 2 * half + i2;
            double rm[R], rs[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { rm[rr] = rT[(size_t)(8 * bi + i) * RS + lane + 32 * rr]; rs[rr] = 0.0; }
#pragma unroll
            for (int k0 = 0; k0 < 8; k0 += KU) {
              // no branch for the pairs below the diagonal of a diagonal block: their cotangent is 0, so g = 0;
              // a branch here splits the unrolled code into basic blocks and stops ptxas overlapping pair steps.
              // KU pair steps advance in lockstep: KU * 2 R independent FMA chains and KU * R exp chains.
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
                double* gdst = gtw + (size_t)(i2 * 8 + k0 + u) * RS + lane;
#pragma unroll
                for (int rr = 0; rr < R; ++rr) gdst[32 * rr] = g[u * R + rr];
                // (row / column sums before the dv FMAs: 100.2 ms; after them: 100.8 ms; g stored after both: 105.3 ms)
#pragma unroll
                for (int rr = 0; rr < R; ++rr) { rs[rr] += g[u * R + rr]; cs[k0 + u][rr] += g[u * R + rr]; }
#pragma unroll
                for (int q = 0; q < QP; ++q)
#pragma unroll
                  for (int rr = 0; rr < R; ++rr) dv[rr][q] = fma(g[u * R + rr], dq[u][q], dv[rr][q]);
              }
            }
#pragma unroll
            for (int rr = 0; rr < R; ++rr) drT[(size_t)(8 * bi + i) * RS + lane + 32 * rr] += rs[rr];
          }
          __syncwarp();
          // ---- phase 2: lane <-> (two pairs, q half, half of the rows); dD accumulators private, rows sequential.
          //      The second row half walks its rows shifted by one so that the two halves hit different banks.
          {
            constexpr int HR = ROWS / 2;
            double acc0[QH], acc1[QH];
#pragma unroll
            for (int j = 0; j < QH; ++j) { acc0[j] = 0.0; acc1[j] = 0.0; }
            double old0[QH], old1[QH];
            if (kSliceRmw && p2_rh == 0) {                  // issued before the row loop so that the L2 latency is hidden
              const double* src = slot + (size_t)(half * PB + 2 * p2_pp) * QP + p2_qh * QH;
#pragma unroll
              for (int j = 0; j < QH; ++j) { old0[j] = __ldcg(src + j); old1[j] = __ldcg(src + QP + j); }
            }
            const double* gp0 = gtw + (size_t)(2 * p2_pp) * RS + p2_rh * HR;
            const double* gp1 = gp0 + RS;
            const double* vp = vt + (size_t)(p2_rh * HR) * 2 * QHP + p2_qh * QHP;
DPGP_UNROLL(DPGP_XP_P2_UNROLL)
            for (int rw = 0; rw < HR; ++rw) {
              const int row = (rw + p2_rh) & (HR - 1);
              const double g0 = gp0[row], g1 = gp1[row];
              double vv[QHP];
#pragma unroll
              for (int j = 0; j < QHP; j += 2) { const double2 t2 = *reinterpret_cast<const double2*>(vp + (size_t)row * 2 * QHP + j); vv[j] = t2.x; vv[j + 1] = t2.y; }
#pragma unroll
              for (int j = 0; j < QH; ++j) { acc0[j] = fma(g0, vv[j], acc0[j]); acc1[j] = fma(g1, vv[j], acc1[j]); }
            }
            if constexpr (DZ) {
              // the two row halves are merged so that lane (rh, pp, qh) owns pair 2 pp + rh of the step: i2 = pp >> 2,
              // k = 2 (pp & 3) + rh; t = (z_m - z_m') dD (the factor 2 is applied by dz_fused_reduce_kernel)
              const int i2 = p2_pp >> 2, kk = 2 * (p2_pp & 3) + p2_rh;
              const int mrow = 8 * bi + 2 * half + i2, mcol = 8 * bj + kk;
              const double* zr = zs + (size_t)mrow * QP + p2_qh * QH;
              const double* zc = zs + (size_t)mcol * QP + p2_qh * QH;
              double t[QH];
#pragma unroll
              for (int j = 0; j < QH; ++j) {
                const double send = p2_rh ? acc0[j] : acc1[j], keepv = p2_rh ? acc1[j] : acc0[j];
                const double mine = keepv + __shfl_xor_sync(0xffffffffu, send, 16);
                t[j] = (zr[j] - zc[j]) * mine;
              }
              // all shuffles first, in converged code; the lane-predicated reductions follow (interleaving them makes
              // ptxas wrap every shuffle in WARPSYNC / ENDCOLLECTIVE: 114 ms instead of 101)
              double rsum[QH], csum[QH];
#pragma unroll
              for (int j = 0; j < QH; ++j) {
                double a = t[j] + __shfl_xor_sync(0xffffffffu, t[j], 16);         // over the 8 columns of the block row
                a += __shfl_xor_sync(0xffffffffu, a, 2);
                rsum[j] = a + __shfl_xor_sync(0xffffffffu, a, 4);
                csum[j] = t[j] + __shfl_xor_sync(0xffffffffu, t[j], 8);          // over the two block rows of the step
              }
              if (p2_rh == 0 && (p2_pp & 3) == 0) {
                double* dst = dzr + (size_t)mrow * QP + p2_qh * QH;
#pragma unroll
                for (int j = 0; j < QH; ++j) red_add_f64_keep(dst + j, rsum[j], keep);
              }
              if (i2 == 0) {
                double* dst = dzc + (size_t)mcol * QP + p2_qh * QH;
#pragma unroll
                for (int j = 0; j < QH; ++j) red_add_f64_keep(dst + j, csum[j], keep);
              }
            } else {
  #pragma unroll
              for (int j = 0; j < QH; ++j) {
                acc0[j] += __shfl_down_sync(0xffffffffu, acc0[j], 16);
                acc1[j] += __shfl_down_sync(0xffffffffu, acc1[j], 16);
              }
              if (p2_rh == 0) {
                double* dst = slot + (size_t)(half * PB + 2 * p2_pp) * QP + p2_qh * QH;
                if (kSliceRmw) {
                  // plain read-modify-write through L2 (the slice address is private to this lane); experiment, see kSliceRmw
  #pragma unroll
                  for (int j = 0; j < QH; ++j) { __stcg(dst + j, old0[j] + acc0[j]); __stcg(dst + QP + j, old1[j] + acc1[j]); }
                } else {
  #pragma unroll
                  for (int j = 0; j < QH; ++j) { red_add_f64_keep(dst + j, acc0[j], keep); red_add_f64_keep(dst + QP + j, acc1[j], keep); }
                }
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
      if (TEAMS == 1) __syncthreads(); else team_barrier(team, kFusedWarps * 32);
      wc[0] = wn[0]; wc[1] = wn[1];
    }
    if (TEAMS > 1) __syncthreads();
    // ---- drain: dv summed over the warps in fixed order, dr transposed back to [row][Mp]
#pragma unroll
    for (int q = 0; q < QP; ++q)
#pragma unroll
      for (int rr = 0; rr < R; ++rr) xdv[((size_t)warp * QP + q) * RS + lane + 32 * rr] = dv[rr][q];
    __syncthreads();
    for (int idx = tid; idx < nc * QP; idx += T) {
      const int row = idx / QP, q = idx - row * QP;
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) a += xdv[((size_t)w * QP + q) * RS + row];
      __stcs(p.dv + ((int64_t)b * p.n + n0) * QP + idx, a);
    }
    {
      double* dst = p.dr + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < nc * p.mp; idx += T) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        double a = drT0[(size_t)m * RS + row];
#pragma unroll
        for (int tm = 1; tm < TEAMS; ++tm) a += drT0[((size_t)tm * p.mp + m) * RS + row];
        __stcs(dst + idx, a);
      }
    }
  }
}

// Sum of the per-CTA dD slices into dDsym [B,M,M,QP] (both triangles; diagonal untouched = 0), fixed order.
struct DdFusedReduceParams {
  const double* part; const int* tags; const unsigned short* sched; double* ddsym;
  int grid, nseg, nrounds, m, b, qp; int64_t ngroups;
};
static __global__ void dd_fused_reduce_kernel(DdFusedReduceParams p) {
  const int64_t per_b = (int64_t)p.nrounds * kFusedWarps * 64 * p.qp;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (b, round, warp, pair, q)
  if (idx >= per_b * p.b) return;
  const int b = (int)(idx / per_b);
  const int64_t rem = idx % per_b;
  const int q = (int)(rem % p.qp);
  const int pr = (int)((rem / p.qp) & 63), slot = (int)(rem / ((int64_t)64 * p.qp));
  const unsigned short it = p.sched[slot];
  if (it == kSchedIdle) return;
  const int m = 8 * (it >> 8) + (pr >> 3), c = 8 * (it & 255) + (pr & 7);
  if (m >= p.m || c >= p.m || m >= c) return;
  // CTA c owns items [items c / grid, items (c+1) / grid) of the (cluster, group) list: only these can hold cluster b
  const int64_t items = p.ngroups * p.b;
  int c_lo = (int)(((int64_t)b * p.ngroups * p.grid) / items) - 1, c_hi = (int)((((int64_t)b + 1) * p.ngroups * p.grid + items - 1) / items) + 1;
  c_lo = max(c_lo, 0); c_hi = min(c_hi, p.grid - 1);
  double s = 0;
  for (int cta = c_lo; cta <= c_hi; ++cta)
    for (int sg = 0; sg < p.nseg; ++sg) {
      const int k = cta * p.nseg + sg;
      if (p.tags[k] == b) s += p.part[(size_t)k * per_b + rem];
    }
  p.ddsym[(((size_t)b * p.m + m) * p.m + c) * p.qp + q] = s;
  p.ddsym[(((size_t)b * p.m + c) * p.m + m) * p.qp + q] = s;
}


// DZ variant: dz[m][q] = 2 sum over (CTA, warp) of (row-side slice - column-side slice), one warp per output, fixed order.
// The total goes to cluster 0 of dzd [B,M,Q] (the layout zchain_kernel fills in the other variants); the rest is zeroed.
struct DzFusedReduceParams { const double* part; double* dzd; int nslices, m, mp, q, qp, b; };
static __global__ void dz_fused_reduce_kernel(DzFusedReduceParams p) {
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (int64_t)p.b * p.m * p.q) return;
  if (gw >= (int64_t)p.m * p.q) { if (lane == 0) p.dzd[gw] = 0.0; return; }
  const int m = (int)(gw / p.q), q = (int)(gw % p.q);
  const size_t one = (size_t)p.mp * p.qp, off = (size_t)m * p.qp + q;
  double s = 0.0;
  for (int k = lane; k < p.nslices; k += 32) {
    const double* base = p.part + (size_t)k * 2 * one + off;
    s += base[0] - base[one];
  }
  s = warp_sum(s);
  if (lane == 0) p.dzd[gw] = 2.0 * s;
}

}  // namespace dpgp
