// psi2 statistic of the RBF-ARD kernel under Gaussian q(X): forward (reference
// src/kernels/rbf_kernel.py:164-199) and its hand-written backward (replaces tf.gradients).
//
// Factorisation used by every kernel here (DESIGN.md "psi2"): with w = g/(2 g s + 1) (g = gamma_bq),
//   log psi2_n[m,m'] = r_nm + r_nm' + sum_q v_nq (z_mq - z_m'q)^2
//   r_nm = 1/2 c_n - 1/2 sum_q w_nq (mu_nq - z_mq)^2,   c_n = 2 log alpha - 1/2 sum_q log(2 g s_nq + 1)
//   v_nq = 1/4 (w_nq - g) = -1/2 g^2 s_nq / (2 g s_nq + 1)
// which is algebraically identical to the reference's
//   2 log alpha - sum_q [ 1/2 log(2 g s + 1) + 1/4 g (z_m - z_m')^2 + g (mu - zbar)^2 / (2 g s + 1) ]
// because (mu - zbar)^2 = 1/2 (mu - z_m)^2 + 1/2 (mu - z_m')^2 - 1/4 (z_m - z_m')^2.
// One "unit" (cluster b, row n, pair m <= m') then costs Q FMAs + 1 add + one exp instead of 3Q + exp.
// r [B,N,Mp] and v [B,N,QP] are produced once per evaluation by prep_rows_kernel and shared by the
// forward and both backward kernels.
#pragma once
#include "common.cuh"

namespace dpgp {

// ------------------------------------------------------------------------------------------------ prep
struct PrepParams {
  const double* mu; const double* s; const double* z; const double* gamma; const double* alpha;
  double* r; double* v;
  int64_t n; int q, m, mp, b; int64_t nchunks;
};
constexpr int kPrepRows = 32;

template <int QP>
__global__ void __launch_bounds__(256) prep_rows_kernel(PrepParams p) {
  __shared__ double ws[kPrepRows][QP], mus[kPrepRows][QP], lden[kPrepRows][QP], cn[kPrepRows];
  const int64_t items = p.nchunks * p.b;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * kPrepRows;
    const int nc = (int)min((int64_t)kPrepRows, p.n - n0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kPrepRows * QP; idx += blockDim.x) {
      int n = idx / QP, q = idx % QP;
      double w = 0, mu = 0, ld = 0, v = 0;
      if (n < nc && q < p.q) {
        double g = p.gamma[b * p.q + q];
        double s = p.s[(n0 + n) * p.q + q];
        mu = p.mu[(n0 + n) * p.q + q];
        double den = fma(2.0 * g, s, 1.0);
        w = g / den;
        v = -0.5 * g * g * s / den;
        ld = log(den);
      }
      ws[n][q] = w; mus[n][q] = mu; lden[n][q] = ld;
      if (n < nc) p.v[((int64_t)b * p.n + n0 + n) * QP + q] = v;
    }
    __syncthreads();
    if (threadIdx.x < kPrepRows) {
      double a = 0;
#pragma unroll
      for (int q = 0; q < QP; ++q) a += lden[threadIdx.x][q];
      cn[threadIdx.x] = 2.0 * log(p.alpha[b]) - 0.5 * a;
    }
    __syncthreads();
    for (int m = threadIdx.x; m < p.mp; m += blockDim.x) {
      double zm[QP];
#pragma unroll
      for (int q = 0; q < QP; ++q) zm[q] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0;
      for (int n = 0; n < nc; ++n) {
        double a = 0;
#pragma unroll
        for (int q = 0; q < QP; ++q) { double d = mus[n][q] - zm[q]; a = fma(ws[n][q] * d, d, a); }
        double r = 0.5 * (cn[n] - a);
        r = fmax(r, kRClamp);
        if (m >= p.m) r = 0.0;
        p.r[((int64_t)b * p.n + n0 + n) * p.mp + m] = r;
      }
    }
  }
}

// --------------------------------------------------------------------------------------------- forward
struct Psi2FwdParams {
  const double* r; const double* v; const double* z;
  double* part;          // [grid*nseg][npass*nthreads*4]
  int* tags;             // [grid*nseg] cluster index of each partial slot, -1 = unused
  int64_t n; int q, m, mp, mt, b, t2, npass, chunk, nseg; int64_t nchunks;
};

// Dynamic shared memory layout (doubles): acc[npass*T*4] | rbuf[2][chunk*mp] | vbuf[2][chunk*QP] | zs[2*mt*QP]
template <int QP, int EXPV>
__global__ void __launch_bounds__(448, 1) psi2_fwd_kernel(Psi2FwdParams p) {
  extern __shared__ __align__(16) double sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  double* acc = sm;
  double* rbuf = acc + (size_t)p.npass * T * 4;
  double* vbuf = rbuf + 2 * (size_t)p.chunk * p.mp;
  double* zs = vbuf + 2 * (size_t)p.chunk * QP;
  Exp<EXPV> ex; ex.init();

  for (int i = tid; i < 2 * p.mt * QP; i += T) {
    int m = i / QP, q = i % QP;
    zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0;
  }
  const int64_t items = p.nchunks * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  if (lo >= hi) return;

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

  // tile of this thread in each pass is fixed for the whole kernel
  int cur_b = -1, seg = 0;
  issue(lo, 0);
  for (int64_t item = lo; item < hi; ++item) {
    const int buf = (int)((item - lo) & 1);
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * p.chunk;
    const int nc = (int)min((int64_t)p.chunk, p.n - n0);
    if (item + 1 < hi) { issue(item + 1, buf ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    if (b != cur_b) {
      if (cur_b >= 0) {      // flush finished cluster
        __syncthreads();
        double* dst = p.part + ((size_t)blockIdx.x * p.nseg + seg) * p.npass * T * 4;
        for (int i = tid; i < p.npass * T * 4; i += T) dst[i] = acc[i];
        if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = cur_b;
        ++seg;
      }
      __syncthreads();
      for (int i = tid; i < p.npass * T * 4; i += T) acc[i] = 0.0;
      cur_b = b;
    }
    __syncthreads();     // tile `buf` (and zs / acc) visible to all threads
    const double* rt = rbuf + (size_t)buf * p.chunk * p.mp;
    const double* vt = vbuf + (size_t)buf * p.chunk * QP;
    for (int pass = 0; pass < p.npass; ++pass) {
      int t = pass * T + tid;
      int ti, tj;
      tile_from_index(t < p.t2 ? t : 0, p.mt, ti, tj);
      const int m0 = 2 * ti, c0 = 2 * tj;
      double d00[QP], d01[QP], d10[QP], d11[QP];
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        double za = zs[m0 * QP + q], zb = zs[(m0 + 1) * QP + q], zc = zs[c0 * QP + q], zd = zs[(c0 + 1) * QP + q];
        double x;
        x = za - zc; d00[q] = x * x;
        x = za - zd; d01[q] = x * x;
        x = zb - zc; d10[q] = x * x;
        x = zb - zd; d11[q] = x * x;
      }
      double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
#pragma unroll 2
      for (int n = 0; n < nc; ++n) {
        const double2 ra = *reinterpret_cast<const double2*>(rt + n * p.mp + m0);
        const double2 rc = *reinterpret_cast<const double2*>(rt + n * p.mp + c0);
        double vq[QP];
#pragma unroll
        for (int q = 0; q < QP; q += 2) {
          const double2 t2 = *reinterpret_cast<const double2*>(vt + n * QP + q);
          vq[q] = t2.x; vq[q + 1] = t2.y;
        }
        double e00 = ra.x + rc.x, e01 = ra.x + rc.y, e10 = ra.y + rc.x, e11 = ra.y + rc.y;
#pragma unroll
        for (int q = 0; q < QP; ++q) {
          e00 = fma(vq[q], d00[q], e00);
          e01 = fma(vq[q], d01[q], e01);
          e10 = fma(vq[q], d10[q], e10);
          e11 = fma(vq[q], d11[q], e11);
        }
        a00 = ex.acc(e00, a00);
        a01 = ex.acc(e01, a01);
        a10 = ex.acc(e10, a10);
        a11 = ex.acc(e11, a11);
      }
      double* a = acc + ((size_t)pass * T + tid) * 4;
      a[0] += a00; a[1] += a01; a[2] += a10; a[3] += a11;
    }
    __syncthreads();     // everyone done with tile `buf` before it is refilled two iterations later
  }
  {
    double* dst = p.part + ((size_t)blockIdx.x * p.nseg + seg) * p.npass * T * 4;
    for (int i = tid; i < p.npass * T * 4; i += T) dst[i] = acc[i];
    if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = cur_b;
  }
}

// Deterministic reduction of the per-CTA partials (fixed slot order) into the symmetric Psi2 [B,M,M].
struct Psi2ReduceParams {
  const double* part; const int* tags; double* psi2;
  int nslots, slot_len, m, mt, t2, b;
};
static __global__ void psi2_reduce_kernel(Psi2ReduceParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // (b, tile, e)
  if (idx >= p.b * p.t2 * 4) return;
  const int b = idx / (p.t2 * 4), rem = idx % (p.t2 * 4), t = rem >> 2, e = rem & 3;
  int ti, tj; tile_from_index(t, p.mt, ti, tj);
  const int m = 2 * ti + (e >> 1), c = 2 * tj + (e & 1);
  if (m >= p.m || c >= p.m || m > c) return;
  double s = 0;
  for (int k = 0; k < p.nslots; ++k)
    if (p.tags[k] == b) s += p.part[(size_t)k * p.slot_len + rem];
  p.psi2[((size_t)b * p.m + m) * p.m + c] = s;
  p.psi2[((size_t)b * p.m + c) * p.m + m] = s;
}

}  // namespace dpgp
