// Backward of the psi2 statistic (replaces TensorFlow autodiff through src/kernels/rbf_kernel.py:164-199).
//
// With E(n,p) = r_nm + r_nm' + sum_q v_nq D_pq (see psi2.cuh), Gs_p the symmetrised cotangent of Psi2 and
// g_np = Gs_p exp(E(n,p)):
//    d r_nm  = sum_{m'} g_n(m,m')          (row + column sums of the symmetric g matrix)   [B,N,Mp]
//    d v_nq  = sum_p g_np D_pq                                                            [B,N,QP]
//    d D_pq  = sum_n g_np v_nq                                                            [B,M,M,QP]
// Two kernels, each with the mapping that makes ITS reduction thread-private:
//   psi2_bwd_n_kernel    lane <-> row n, sweeps the pair triangle in 8x8 blocks: d r, d v private to a lane
//   psi2_bwd_pair_kernel thread <-> pairs (as the forward), loops over n: d D private to a thread
// The chain from (d r, d v, d D) to (mu, s, Z, gamma, alpha) is in chain.cuh.
#pragma once
#include "../common.cuh"
#include "../psi2.cuh"
#include "../psi2_bwd_fused.cuh"   // sym_cotangent

namespace dpgp {

// ------------------------------------------------------------------------------------------ pair side
struct Psi2BwdPairParams {
  const double* r; const double* v; const double* z; const double* gbar;   // gbar: d/dPsi2 [B,M,M]
  const double* exptab;
  double* part;          // [grid*nseg][TC*2*QP]   (TC = consumer threads)
  int* tags;             // [grid*nseg]
  int64_t n; int q, m, mp, mt, b, t2, jb, ng, chunk, nseg; int64_t nchunks;
};

// Same producer/consumer ring as psi2_fwd_kernel (see psi2.cuh).  A consumer thread owns ONE half tile
// (row m, columns c0, c0+1) for the whole kernel -- D and the dD accumulators (4 QP doubles) stay in
// registers -- so the pairs are split over `jb` CTAs; the jb CTAs of a group walk the same (cluster, chunk)
// items at the same time and share the r / v tiles through L2.
// smem: stage[kStages][chunk*(mp+QP)] f64 | zs[2*mt*QP] f64 | full[kStages], empty[kStages] u64
template <int QP, int EXPV>
__global__ void __launch_bounds__(384, 1) psi2_bwd_pair_kernel(Psi2BwdPairParams p) {
  extern __shared__ __align__(16) double sm[];
  const int T = blockDim.x, tid = threadIdx.x, TC = T - 32, ncw = TC / 32;
  double* stage = sm;
  const size_t stage_len = (size_t)p.chunk * (p.mp + QP);
  double* zs = stage + kStages * stage_len;
  uint64_t* full = reinterpret_cast<uint64_t*>(zs + 2 * p.mt * QP);
  uint64_t* empty = full + kStages;
  double* etab = reinterpret_cast<double*>(empty + kStages);
  if (EXPV >= 4) load_exp_table(etab, p.exptab);
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  const int j = blockIdx.x % p.jb, grp = blockIdx.x / p.jb;
  if (grp >= p.ng) return;
  for (int i = tid; i < 2 * p.mt * QP; i += T) {
    int m = i / QP, q = i % QP;
    zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0;
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncw); }
    mbar_fence_init();
  }
  __syncthreads();
  const int64_t items = p.nchunks * p.b;
  const int64_t lo = items * grp / p.ng, hi = items * (grp + 1) / p.ng;
  if (lo >= hi) return;

  if (tid >= TC) {
    if (tid == TC) {
      for (int64_t item = lo; item < hi; ++item) {
        const int k = (int)(item - lo), s = k % kStages;
        if (k >= kStages) mbar_wait(&empty[s], ((k / kStages) - 1) & 1);
        const int b = (int)(item / p.nchunks);
        const int64_t n0 = (item % p.nchunks) * p.chunk;
        const int nc = (int)min((int64_t)p.chunk, p.n - n0);
        double* rd = stage + s * stage_len;
        double* vd = rd + (size_t)p.chunk * p.mp;
        const unsigned rbytes = (unsigned)(nc * p.mp * 8), vbytes = (unsigned)(nc * QP * 8);
        mbar_expect_tx(&full[s], rbytes + vbytes);
        bulk_g2s(rd, p.r + ((int64_t)b * p.n + n0) * p.mp, rbytes, &full[s]);
        bulk_g2s(vd, p.v + ((int64_t)b * p.n + n0) * QP, vbytes, &full[s]);
      }
    }
    return;
  }
  Exp<EXPV> ex; ex.init(etab);
  const int h = j * TC + tid;
  const bool valid = h < 2 * p.t2;
  int ti, tj; tile_from_index(valid ? (h >> 1) : 0, p.mt, ti, tj);
  const int m = 2 * ti + (h & 1), c0 = 2 * tj;
  double d0[QP], d1[QP], g0[QP], g1[QP];
#pragma unroll
  for (int q = 0; q < QP; ++q) {
    double x = zs[m * QP + q] - zs[c0 * QP + q]; d0[q] = x * x;
    x = zs[m * QP + q] - zs[(c0 + 1) * QP + q]; d1[q] = x * x;
    g0[q] = 0; g1[q] = 0;
  }
  auto flush = [&](int seg, int b) {
    double* dst = p.part + (((size_t)blockIdx.x * p.nseg + seg) * TC + tid) * 2 * QP;
#pragma unroll
    for (int q = 0; q < QP; ++q) { dst[q] = g0[q]; dst[QP + q] = g1[q]; g0[q] = 0; g1[q] = 0; }
    if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = b;
  };
  auto weights = [&](int b, double& w0, double& w1) {
    const double* gb = p.gbar + (size_t)b * p.m * p.m;
    w0 = valid ? sym_cotangent(gb, m, c0, p.m) : 0.0;
    w1 = valid ? sym_cotangent(gb, m, c0 + 1, p.m) : 0.0;
  };
  int cur_b = (int)(lo / p.nchunks), seg = 0;
  double w0, w1; weights(cur_b, w0, w1);
  for (int64_t item = lo; item < hi; ++item) {
    const int k = (int)(item - lo), s = k % kStages;
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * p.chunk;
    const int nc = (int)min((int64_t)p.chunk, p.n - n0);
    if (b != cur_b) { flush(seg++, cur_b); cur_b = b; weights(b, w0, w1); }
    mbar_wait(&full[s], (k / kStages) & 1);
    const double* rt = stage + s * stage_len;
    const double* vt = rt + (size_t)p.chunk * p.mp;
#pragma unroll 1
    for (int n = 0; n < nc; n += 2) {
      const int n1 = (n + 1 < nc) ? n + 1 : n;             // odd tail: second row repeats the first with weight 0
      const double wt = (n + 1 < nc) ? 1.0 : 0.0;
      const double ra0 = rt[n * p.mp + m], ra1 = rt[n1 * p.mp + m];
      const double2 rc0 = *reinterpret_cast<const double2*>(rt + n * p.mp + c0);
      const double2 rc1 = *reinterpret_cast<const double2*>(rt + n1 * p.mp + c0);
      double va[QP], vb[QP];
#pragma unroll
      for (int q = 0; q < QP; q += 2) {
        const double2 t0 = *reinterpret_cast<const double2*>(vt + n * QP + q);
        const double2 t1 = *reinterpret_cast<const double2*>(vt + n1 * QP + q);
        va[q] = t0.x; va[q + 1] = t0.y; vb[q] = t1.x; vb[q + 1] = t1.y;
      }
      double e[4] = {ra0 + rc0.x, ra0 + rc0.y, ra1 + rc1.x, ra1 + rc1.y};
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        e[0] = fma(va[q], d0[q], e[0]);
        e[1] = fma(va[q], d1[q], e[1]);
        e[2] = fma(vb[q], d0[q], e[2]);
        e[3] = fma(vb[q], d1[q], e[3]);
      }
      const double w[4] = {w0, w1, w0 * wt, w1 * wt};
      double x[4];
      exp_scaled_k<EXPV, 4>(ex, e, w, x);
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        g0[q] = fma(x[0], va[q], g0[q]);
        g1[q] = fma(x[1], va[q], g1[q]);
      }
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        g0[q] = fma(x[2], vb[q], g0[q]);
        g1[q] = fma(x[3], vb[q], g1[q]);
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
  flush(seg, cur_b);
}

// dD partials -> dDsym [B,M,M,QP] (both triangles, diagonal zero), fixed summation order.
struct DdReduceParams {
  const double* part; const int* tags; double* ddsym;
  int grid, jb, T, m, mt, t2, b, qp, nseg;
};
static __global__ void dd_reduce_kernel(DdReduceParams p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // (b, halftile, e, q)
  const int64_t per_b = (int64_t)2 * p.t2 * 2 * p.qp;
  if (idx >= per_b * p.b) return;
  const int b = (int)(idx / per_b);
  int rem = (int)(idx % per_b);
  const int q = rem % p.qp; rem /= p.qp;
  const int e = rem & 1, h = rem >> 1;
  int ti, tj; tile_from_index(h >> 1, p.mt, ti, tj);
  const int m = 2 * ti + (h & 1), c = 2 * tj + e;
  if (m >= p.m || c >= p.m || m >= c) return;
  const int j = h / p.T, tid = h % p.T;
  double s = 0;
  for (int cta = j; cta < p.grid; cta += p.jb)
    for (int seg = 0; seg < p.nseg; ++seg)
      if (p.tags[cta * p.nseg + seg] == b)
        s += p.part[((((size_t)cta * p.nseg + seg) * p.T + tid) * 2 + e) * p.qp + q];
  p.ddsym[(((size_t)b * p.m + m) * p.m + c) * p.qp + q] = s;
  p.ddsym[(((size_t)b * p.m + c) * p.m + m) * p.qp + q] = s;
}

// --------------------------------------------------------------------------------------------- n side
// Block tables shared by every row n.  The pair triangle is swept in blocks of 8 rows x 4 columns
// (block-row bi = rows 8bi..8bi+7, column blocks bj4 >= 2 bi); for row i of block blk and column k:
//   dtab[((blk*8+i)*4+k)*QP + q] = (z_row - z_col)_q^2          (depends on Z only)
//   gtab[b][(blk*8+i)*4+k]      = symmetrised cotangent of Psi2[b], 0 below the diagonal / outside M
// They live in global memory (0.7 MB + 70 KB per cluster at M = 128: L1/L2 resident) and are read with
// warp-uniform addresses, so the n-side kernel needs no shared-memory staging and no CTA barrier.
__host__ __device__ inline int nside_num_blocks(int mp) {
  const int nb8 = mp / 8, nb4 = mp / 4;
  int n = 0;
  for (int bi = 0; bi < nb8; ++bi) n += nb4 - 2 * bi;
  return n;
}
struct BlockTabParams { const double* z; const double* gbar; double* dtab; double* gtab; int q, qp, m, mp, b; };
static __global__ void block_tables_kernel(BlockTabParams p) {
  const int nb4 = p.mp / 4, nblk = nside_num_blocks(p.mp);
  const int64_t total = (int64_t)nblk * 32;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int u = (int)(idx & 31); int blk = (int)(idx >> 5);
    int bi = 0, rem = blk;
    while (rem >= nb4 - 2 * bi) { rem -= nb4 - 2 * bi; ++bi; }
    const int bj4 = 2 * bi + rem;
    const int m = 8 * bi + (u >> 2), c = 4 * bj4 + (u & 3);
    for (int q = 0; q < p.qp; ++q) {
      double x = 0.0;
      if (m < p.m && c < p.m && q < p.q) x = p.z[m * p.q + q] - p.z[c * p.q + q];
      p.dtab[idx * p.qp + q] = x * x;
    }
    for (int b = 0; b < p.b; ++b)
      p.gtab[(int64_t)b * total + idx] = sym_cotangent(p.gbar + (size_t)b * p.m * p.m, m, c, p.m);
  }
}

struct Psi2BwdNParams {
  const double* r; const double* v; const double* dtab; const double* gtab; const double* exptab;
  double* dr;            // [B,N,Mp]  (may alias r: a lane overwrites only its own row after it has finished reading it)
  double* dv;            // [B,N,QP]
  int64_t n; int q, m, mp, b; int64_t ngroups;     // groups of blockDim.x rows
};

// lane <-> row n.  smem (doubles): dracc[mp][T] -- lane-private accumulators of d r_nm; no barrier in the kernel.
// The 4 units of one row step (row i, columns k = 0..3) advance in lockstep: 4 independent DFMA chains.
template <int QP, int EXPV>
__global__ void __launch_bounds__(192, 1) psi2_bwd_n_kernel(Psi2BwdNParams p) {
  extern __shared__ __align__(16) double sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  double* dracc = sm;
  double* etab = sm + (size_t)p.mp * T;
  if (EXPV >= 4) { load_exp_table(etab, p.exptab); __syncthreads(); }
  Exp<EXPV> ex; ex.init(etab);
  const int nb8 = p.mp / 8, nb4 = p.mp / 4, nblk = nside_num_blocks(p.mp);
  const int64_t items = p.ngroups * p.b;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / p.ngroups);
    const int64_t n = (item % p.ngroups) * T + tid;
    const bool live = n < p.n;
    const int64_t nn = live ? n : p.n - 1;           // clamp: dead lanes compute on a valid row, never store
    const double* rrow = p.r + ((int64_t)b * p.n + nn) * p.mp;
    const double* gt_b = p.gtab + (int64_t)b * nblk * 32;
    double vq[QP], dv[QP];
#pragma unroll
    for (int q = 0; q < QP; ++q) { vq[q] = p.v[((int64_t)b * p.n + nn) * QP + q]; dv[q] = 0; }
    for (int i = 0; i < p.mp; ++i) dracc[(size_t)i * T + tid] = 0.0;
    int blk = 0;
    for (int bi = 0; bi < nb8; ++bi) {
      double rr[8], rs[8];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(rrow + 8 * bi + i);
        rr[i] = t2.x; rr[i + 1] = t2.y; rs[i] = 0; rs[i + 1] = 0;
      }
      for (int bj4 = 2 * bi; bj4 < nb4; ++bj4, ++blk) {
        double rc[4], cs[4] = {0, 0, 0, 0};
        {
          const double2 t0 = *reinterpret_cast<const double2*>(rrow + 4 * bj4);
          const double2 t1 = *reinterpret_cast<const double2*>(rrow + 4 * bj4 + 2);
          rc[0] = t0.x; rc[1] = t0.y; rc[2] = t1.x; rc[3] = t1.y;
        }
        const int imax = (bj4 == 2 * bi) ? 4 : 8;    // rows 4..7 of the first diagonal block lie below the diagonal
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (i < imax) {                            // warp-uniform
            const double* dt = p.dtab + ((size_t)blk * 8 + i) * 4 * QP;
            const double2* gp = reinterpret_cast<const double2*>(gt_b + ((size_t)blk * 8 + i) * 4);
            double dq[4][QP];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int q = 0; q < QP; q += 2) {
                const double2 t2 = __ldg(reinterpret_cast<const double2*>(dt + k * QP + q));
                dq[k][q] = t2.x; dq[k][q + 1] = t2.y;
              }
            const double2 ga = __ldg(gp), gb = __ldg(gp + 1);
            const double w[4] = {ga.x, ga.y, gb.x, gb.y};
            double e[4] = {rr[i] + rc[0], rr[i] + rc[1], rr[i] + rc[2], rr[i] + rc[3]};
#pragma unroll
            for (int q = 0; q < QP; ++q) {
#pragma unroll
              for (int k = 0; k < 4; ++k) e[k] = fma(vq[q], dq[k][q], e[k]);
            }
            double g[4];
            exp_scaled_k<EXPV, 4>(ex, e, w, g);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int q = 0; q < QP; ++q) dv[q] = fma(g[k], dq[k][q], dv[q]);
            }
            rs[i] += (g[0] + g[1]) + (g[2] + g[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) cs[k] += g[k];
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) dracc[(size_t)(4 * bj4 + k) * T + tid] += cs[k];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) dracc[(size_t)(8 * bi + i) * T + tid] += rs[i];
    }
    if (live) {
      double* drow = p.dr + ((int64_t)b * p.n + n) * p.mp;
      for (int i = 0; i < p.mp; i += 2) {
        double2 t2; t2.x = dracc[(size_t)i * T + tid]; t2.y = dracc[(size_t)(i + 1) * T + tid];
        *reinterpret_cast<double2*>(drow + i) = t2;
      }
#pragma unroll
      for (int q = 0; q < QP; ++q) p.dv[((int64_t)b * p.n + n) * QP + q] = dv[q];
    }
  }
}

}  // namespace dpgp
