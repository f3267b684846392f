// Fused backward of the psi2 statistic: ONE exp per (cluster, row, pair m <= m') unit feeds all three
// reductions (replaces TensorFlow autodiff through src/kernels/rbf_kernel.py:164-199; supersedes the two
// kernels of psi2_bwd.cuh, which evaluate the exponential twice).
//
//   g_np  = Gs_p exp(r_nm + r_nm' + sum_q v_nq D_pq)           (psi2.cuh; Gs = symmetrised cotangent of Psi2)
//   dr_nm = sum_{m'} g_n(m,m')     dv_nq = sum_p g_np D_pq      dD_pq = sum_n g_np v_nq
//
// Mapping.  lane <-> R rows of a group of 32 R rows, so dr and dv are thread-private sums.  The M/8 x M/8
// triangle of 8x8 pair blocks is walked in ROUNDS built on the host from a round-robin tournament
// (1-factorisation of the complete graph on the m-blocks): the <= 8 blocks of a round touch disjoint m-blocks,
// each goes to one of the 8 warps, and all warps work on the same row group.  Hence
//   * dr accumulates in ONE shared [Mp][rows] array without conflicts (a CTA barrier separates rounds),
//     in a fixed order: results are bitwise reproducible;
//   * dv stays in registers over all rounds and is summed over the 8 warps once per row group;
//   * dD needs the only cross-lane reduction: the Q products g v_q of a pair are transposed-and-reduced
//     over the 32 lanes with shuffles (12 DADD per pair step at Q = 10, amortised over R rows) and the
//     totals are added into a per-CTA slice of global memory with red.global.add.f64.  Every address of
//     a slice is only ever updated by one lane of one warp, in program order -> deterministic.  Slices
//     are summed over CTAs by dd_fused_reduce_kernel in fixed order.
// FP64-pipe issues per unit at Q = 10, R = 2: 11 (exponent) + 9 (table exp incl. weight) + 10 (dv) +
// 2 (dr) + 10 (dD products) + 6 (reduction) = 48, against 2 x 26 + 22 = 74 for the two-kernel version.
#pragma once
#include "common.cuh"
#include "psi2_bwd.cuh"

namespace dpgp {

constexpr int kFusedWarps = 8;
constexpr unsigned short kSchedIdle = 0xffff;

struct Psi2BwdFusedParams {
  const double* r; const double* v; const double* z; const double* gbar; const double* exptab;
  const unsigned short* sched;   // [nrounds][8]: (bi << 8) | bj, or kSchedIdle
  double* dr;                    // [B,N,Mp]   may alias r (a CTA stages its rows before it overwrites them)
  double* dv;                    // [B,N,QP]
  double* part;                  // [grid*nseg][nrounds*8*64*QP]   zero on entry
  int* tags;                     // [grid*nseg]
  int64_t n; int q, m, mp, b, nrounds, nseg; int64_t ngroups;
};

// Transposing reduction over the 32 lanes: on entry every lane holds K values x[0..K-1]; on exit x[0] of lane l
// holds the 32-lane total of value `fused_owner_q(l)` (or garbage-free zero/duplicate for non-owners).
template <int K, int OFF>
struct TReduce {
  template <int KMAX>
  static __device__ __forceinline__ void run(double (&x)[KMAX], int lane) {
    if constexpr (K == 1) {
      x[0] += __shfl_xor_sync(0xffffffffu, x[0], OFF);
    } else {
      constexpr int H = (K + 1) / 2;
      const bool up = (lane & OFF) != 0;
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const double lo = x[j], hi = (j + H < K) ? x[j + H] : 0.0;
        const double send = up ? lo : hi, keep = up ? hi : lo;
        x[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
      }
    }
    if constexpr (OFF > 1) TReduce<(K + 1) / 2, OFF / 2>::run(x, lane);
  }
};
// Which of the K reduced values lane `lane` owns after TReduce<K,16> (-1: none).
template <int K>
__host__ __device__ inline int fused_owner_q(int lane) {
  int ks[5]; ks[0] = K;
  for (int i = 1; i < 5; ++i) ks[i] = (ks[i - 1] + 1) / 2;
  int idx = 0;
  for (int level = 4; level >= 0; --level) {
    const int off = 16 >> level, k = ks[level], h = (k + 1) / 2;
    const bool up = (lane & off) != 0;
    if (k == 1) { if (up) return -1; }
    else { if (up) idx += h; if (idx >= k) return -1; }
  }
  return idx;
}

__device__ __forceinline__ void red_add_f64(double* addr, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// smem (doubles): rT[mp*RS] | drT[mp*RS] | zs[mp*QP] | etab[256] | dtab[8][64*(QP+2)]  (dtab aliases xdv[8][QP][RS])
template <int QP, int R>
__host__ __device__ inline size_t fused_smem_bytes(int mp) {
  const int RS = 32 * R + 1;
  const size_t dt = (size_t)kFusedWarps * 64 * (QP + 2), xd = (size_t)kFusedWarps * QP * RS;
  return (2 * (size_t)mp * RS + (size_t)mp * QP + kExpTabSize + (dt > xd ? dt : xd)) * 8;
}

template <int QP, int EXPV, int R>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) psi2_bwd_fused_kernel(Psi2BwdFusedParams p) {
  extern __shared__ __align__(16) double sm[];
  constexpr int RS = 32 * R + 1, ROWS = 32 * R, DS = QP + 2, T = kFusedWarps * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* rT = sm;
  double* drT = rT + (size_t)p.mp * RS;
  double* zs = drT + (size_t)p.mp * RS;
  double* etab = zs + (size_t)p.mp * QP;
  double* dtab = etab + kExpTabSize;
  double* xdv = dtab;                                   // alias, used only between the last round and the next fill
  double* dtw = dtab + (size_t)warp * 64 * DS;

  for (int i = tid; i < p.mp * QP; i += T) { const int m = i / QP, q = i % QP; zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0; }
  load_exp_table(etab, p.exptab);
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  Exp<EXPV> ex; ex.init(etab);
  const int my_q = fused_owner_q<QP>(lane);
  const size_t slice_len = (size_t)p.nrounds * kFusedWarps * 64 * QP;

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
    __syncthreads();                                    // previous group's drain has finished with rT / drT / xdv
    {
      const double* src = p.r + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < ROWS * p.mp; idx += T) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        rT[(size_t)m * RS + row] = (row < nc) ? __ldcs(src + idx) : kRClamp;      // dead rows: exp(2 kRClamp) is exactly 0 in every variant
      }
      for (int idx = tid; idx < p.mp * RS; idx += T) drT[idx] = 0.0;
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
    __syncthreads();

    for (int round = 0; round < p.nrounds; ++round) {
      const unsigned short it = p.sched[round * kFusedWarps + warp];
      if (it != kSchedIdle) {
        const int bi = it >> 8, bj = it & 255;
        const bool diag = (bi == bj);
        // ---- this block's table: D[pair][q] and the symmetrised cotangent
        for (int idx = lane; idx < 64; idx += 32) {
          const int i = idx >> 3, k = idx & 7, m = 8 * bi + i, c = 8 * bj + k;
#pragma unroll
          for (int q = 0; q < QP; ++q) { const double d = zs[m * QP + q] - zs[c * QP + q]; dtw[idx * DS + q] = d * d; }
          dtw[idx * DS + QP] = sym_cotangent(gb, m, c, p.m);
          dtw[idx * DS + QP + 1] = 0.0;
        }
        __syncwarp();
        double* slot = mypart + ((size_t)(round * kFusedWarps + warp) * 64) * QP + (my_q >= 0 ? my_q : 0);
        const double* rcol = rT + (size_t)(8 * bj) * RS + lane;
        double cs[8][R];
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int rr = 0; rr < R; ++rr) cs[k][rr] = 0.0;
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          double rm[R], rs[R];
#pragma unroll
          for (int rr = 0; rr < R; ++rr) { rm[rr] = rT[(size_t)(8 * bi + i) * RS + lane + 32 * rr]; rs[rr] = 0.0; }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (diag && k < i) continue;               // warp-uniform
            const double* dt = dtw + (i * 8 + k) * DS;
            double dq[QP];
#pragma unroll
            for (int q = 0; q < QP; q += 2) { const double2 t2 = *reinterpret_cast<const double2*>(dt + q); dq[q] = t2.x; dq[q + 1] = t2.y; }
            const double wgt = dt[QP];
            double e[R], w[R], g[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { e[rr] = rm[rr] + rcol[(size_t)k * RS + 32 * rr]; w[rr] = wgt; }
#pragma unroll
            for (int q = 0; q < QP; ++q)
#pragma unroll
              for (int rr = 0; rr < R; ++rr) e[rr] = fma(vq[rr][q], dq[q], e[rr]);
            exp_scaled_k<EXPV, R>(ex, e, w, g);
#pragma unroll
            for (int q = 0; q < QP; ++q)
#pragma unroll
              for (int rr = 0; rr < R; ++rr) dv[rr][q] = fma(g[rr], dq[q], dv[rr][q]);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) { rs[rr] += g[rr]; cs[k][rr] += g[rr]; }
            double x[QP];
#pragma unroll
            for (int q = 0; q < QP; ++q) {
              x[q] = g[0] * vq[0][q];
#pragma unroll
              for (int rr = 1; rr < R; ++rr) x[q] = fma(g[rr], vq[rr][q], x[q]);
            }
            TReduce<QP, 16>::run(x, lane);
            if (my_q >= 0) red_add_f64(slot + (i * 8 + k) * QP, x[0]);
          }
#pragma unroll
          for (int rr = 0; rr < R; ++rr) drT[(size_t)(8 * bi + i) * RS + lane + 32 * rr] += rs[rr];
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
          for (int rr = 0; rr < R; ++rr) drT[(size_t)(8 * bj + k) * RS + lane + 32 * rr] += cs[k][rr];
      }
      __syncthreads();
    }
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

}  // namespace dpgp
