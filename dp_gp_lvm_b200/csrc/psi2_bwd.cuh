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
#include "common.cuh"

namespace dpgp {

__device__ __forceinline__ double sym_cotangent(const double* gb, int m, int c, int M) {
  if (m >= M || c >= M || m > c) return 0.0;
  if (m == c) return gb[(size_t)m * M + m];
  return gb[(size_t)m * M + c] + gb[(size_t)c * M + m];
}

// ------------------------------------------------------------------------------------------ pair side
struct Psi2BwdPairParams {
  const double* r; const double* v; const double* z; const double* gbar;   // gbar: d/dPsi2 [B,M,M]
  double* part;          // [grid*nseg][T*2*QP]
  int* tags;             // [grid*nseg]
  int64_t n; int q, m, mp, mt, b, t2, jb, ng, chunk, nseg; int64_t nchunks;
};

// smem (doubles): rbuf[2][chunk*mp] | vbuf[2][chunk*QP] | zs[2*mt*QP]
template <int QP, int EXPV>
__global__ void __launch_bounds__(448, 1) psi2_bwd_pair_kernel(Psi2BwdPairParams p) {
  extern __shared__ __align__(16) double sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  double* rbuf = sm;
  double* vbuf = rbuf + 2 * (size_t)p.chunk * p.mp;
  double* zs = vbuf + 2 * (size_t)p.chunk * QP;
  Exp<EXPV> ex; ex.init();
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  const int j = blockIdx.x % p.jb, grp = blockIdx.x / p.jb;
  if (grp >= p.ng) return;
  for (int i = tid; i < 2 * p.mt * QP; i += T) {
    int m = i / QP, q = i % QP;
    zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0;
  }
  const int64_t items = p.nchunks * p.b;
  const int64_t lo = items * grp / p.ng, hi = items * (grp + 1) / p.ng;
  if (lo >= hi) return;
  __syncthreads();

  const int h = j * T + tid;
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

  auto issue = [&](int64_t item, int buf) {
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * p.chunk;
    const int nc = (int)min((int64_t)p.chunk, p.n - n0);
    const double* rs = p.r + ((int64_t)b * p.n + n0) * p.mp;
    const double* vs = p.v + ((int64_t)b * p.n + n0) * QP;
    double* rd = rbuf + (size_t)buf * p.chunk * p.mp;
    double* vd = vbuf + (size_t)buf * p.chunk * QP;
    for (int i = tid * 2; i < nc * p.mp; i += T * 2) cp_async16(rd + i, rs + i);
    for (int i = tid * 2; i < nc * QP; i += T * 2) cp_async16(vd + i, vs + i);
    cp_async_commit();
  };
  auto flush = [&](int seg, int b) {
    double* dst = p.part + (((size_t)blockIdx.x * p.nseg + seg) * T + tid) * 2 * QP;
#pragma unroll
    for (int q = 0; q < QP; ++q) { dst[q] = g0[q]; dst[QP + q] = g1[q]; g0[q] = 0; g1[q] = 0; }
    if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = b;
  };

  int cur_b = -1, seg = 0;
  double w0 = 0, w1 = 0;
  issue(lo, 0);
  for (int64_t item = lo; item < hi; ++item) {
    const int buf = (int)((item - lo) & 1);
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * p.chunk;
    const int nc = (int)min((int64_t)p.chunk, p.n - n0);
    if (item + 1 < hi) { issue(item + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    if (b != cur_b) {
      if (cur_b >= 0) flush(seg++, cur_b);
      cur_b = b;
      const double* gb = p.gbar + (size_t)b * p.m * p.m;
      w0 = valid ? sym_cotangent(gb, m, c0, p.m) : 0.0;
      w1 = valid ? sym_cotangent(gb, m, c0 + 1, p.m) : 0.0;
    }
    __syncthreads();
    const double* rt = rbuf + (size_t)buf * p.chunk * p.mp;
    const double* vt = vbuf + (size_t)buf * p.chunk * QP;
#pragma unroll 2
    for (int n = 0; n < nc; ++n) {
      const double ra = rt[n * p.mp + m];
      const double2 rc = *reinterpret_cast<const double2*>(rt + n * p.mp + c0);
      double vq[QP];
#pragma unroll
      for (int q = 0; q < QP; q += 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(vt + n * QP + q);
        vq[q] = t2.x; vq[q + 1] = t2.y;
      }
      double e0 = ra + rc.x, e1 = ra + rc.y;
#pragma unroll
      for (int q = 0; q < QP; ++q) { e0 = fma(vq[q], d0[q], e0); e1 = fma(vq[q], d1[q], e1); }
      const double x0 = ex.scaled(e0, w0), x1 = ex.scaled(e1, w1);
#pragma unroll
      for (int q = 0; q < QP; ++q) { g0[q] = fma(x0, vq[q], g0[q]); g1[q] = fma(x1, vq[q], g1[q]); }
    }
    __syncthreads();
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
struct Psi2BwdNParams {
  const double* r; const double* v; const double* z; const double* gbar;
  double* dr;            // [B,N,Mp]  (may alias r: a CTA overwrites only rows it has finished reading)
  double* dv;            // [B,N,QP]
  int64_t n; int q, m, mp, b; int64_t ngroups;     // groups of blockDim.x rows
};

// smem (doubles): dracc[mp][T] | dtab[2][64][QP] | gtab[2][64] | zs[mp][QP]
template <int QP, int EXPV>
__global__ void __launch_bounds__(160, 1) psi2_bwd_n_kernel(Psi2BwdNParams p) {
  extern __shared__ __align__(16) double sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  double* dracc = sm;
  double* dtab = dracc + (size_t)p.mp * T;
  double* gtab = dtab + 2 * 64 * QP;
  double* zs = gtab + 2 * 64;
  Exp<EXPV> ex; ex.init();
  const int nb = p.mp / 8;
  for (int i = tid; i < p.mp * QP; i += T) {
    int m = i / QP, q = i % QP;
    zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0;
  }
  const int64_t items = p.ngroups * p.b;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / p.ngroups);
    const int64_t n = (item % p.ngroups) * T + tid;
    const bool live = n < p.n;
    const int64_t nn = live ? n : p.n - 1;           // clamp: dead lanes compute on a valid row, never store
    const double* rrow = p.r + ((int64_t)b * p.n + nn) * p.mp;
    const double* gb = p.gbar + (size_t)b * p.m * p.m;
    double vq[QP], dv[QP];
#pragma unroll
    for (int q = 0; q < QP; ++q) { vq[q] = p.v[((int64_t)b * p.n + nn) * QP + q]; dv[q] = 0; }
    __syncthreads();                                   // previous item's dracc reads are done
    for (int i = 0; i < p.mp; ++i) dracc[(size_t)i * T + tid] = 0.0;

    auto fill = [&](int bi, int bj, int buf) {       // tables of block (bi,bj): D and symmetrised cotangent
      for (int u = tid; u < 64; u += T) {
        const int m = 8 * bi + (u >> 3), c = 8 * bj + (u & 7);
        gtab[buf * 64 + u] = sym_cotangent(gb, m, c, p.m);
#pragma unroll
        for (int q = 0; q < QP; ++q) {
          double x = zs[m * QP + q] - zs[c * QP + q];
          dtab[(buf * 64 + u) * QP + q] = x * x;
        }
      }
    };
    int buf = 0;
    fill(0, 0, 0);
    __syncthreads();
    for (int bi = 0; bi < nb; ++bi) {
      double rr[8], rs[8];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const double2 t2 = *reinterpret_cast<const double2*>(rrow + 8 * bi + i);
        rr[i] = t2.x; rr[i + 1] = t2.y; rs[i] = 0; rs[i + 1] = 0;
      }
      for (int bj = bi; bj < nb; ++bj) {
        // prefetch next block's tables into the other buffer
        int nbi = bi, nbj = bj + 1;
        if (nbj == nb) { nbi = bi + 1; nbj = nbi; }
        if (nbi < nb) fill(nbi, nbj, buf ^ 1);
        double rc[8], cs[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          const double2 t2 = *reinterpret_cast<const double2*>(rrow + 8 * bj + i);
          rc[i] = t2.x; rc[i + 1] = t2.y; cs[i] = 0; cs[i + 1] = 0;
        }
        const double* dt = dtab + (size_t)buf * 64 * QP;
        const double* gt = gtab + buf * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            if (bi == bj && jj < i) continue;        // block-uniform: no divergence
            const int u = i * 8 + jj;
            double dq[QP];
#pragma unroll
            for (int q = 0; q < QP; q += 2) {
              const double2 t2 = *reinterpret_cast<const double2*>(dt + u * QP + q);
              dq[q] = t2.x; dq[q + 1] = t2.y;
            }
            double e = rr[i] + rc[jj];
#pragma unroll
            for (int q = 0; q < QP; ++q) e = fma(vq[q], dq[q], e);
            const double g = ex.scaled(e, gt[u]);
#pragma unroll
            for (int q = 0; q < QP; ++q) dv[q] = fma(g, dq[q], dv[q]);
            rs[i] += g; cs[jj] += g;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) dracc[(size_t)(8 * bj + jj) * T + tid] += cs[jj];
        __syncthreads();                               // tables of the next block complete / current free
        buf ^= 1;
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
