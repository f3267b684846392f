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
// One persistent CTA per SM: `ncw` consumer warps + one producer warp.  The producer streams [rows x Mp] tiles
// of r and [rows x QP] tiles of v through a ring of `kStages` shared-memory stages with 1-D bulk async copies
// (TMA engine) signalled on mbarriers; consumer warps never synchronise with each other (a CTA-wide barrier
// per tile cost 13 % in the first version: with priority scheduling the last warp runs alone at every barrier).
// Each consumer thread owns one 2x2 tile of (m, m') pairs per pass (D = (z_m - z_m')^2 for its 4 pairs in
// registers) and accumulates sum_n exp(E) privately; per-thread accumulators of all passes live in shared
// memory slots that only their owner touches.
constexpr int kStages = 3;
// Two rows per consumer iteration (8 exponent / exp chains in lockstep): 40.6 -> 39.0 ms at 262 144 rows.
// Unroll of the two-row loop: 2 runs 40.8 ms, 1 (default) 38.9 ms.
#ifndef DPGP_XP_FWD_UNROLL
#define DPGP_XP_FWD_UNROLL 1
#endif
#define DPGP_FWD_PRAGMA_(x) _Pragma(#x)
#define DPGP_FWD_UNROLL(n) DPGP_FWD_PRAGMA_(unroll n)
#ifndef DPGP_FWD_ROWS2
#define DPGP_FWD_ROWS2 1
#endif

struct Psi2FwdParams {
  const double* r; const double* v; const double* z; const double* exptab;
  double* part;          // [grid*nseg][npass*TC*4]
  int* tags;             // [grid*nseg] cluster index of each partial slot, -1 = unused
  int64_t n; int q, m, mp, mt, b, t2, npass, chunk, nseg; int64_t nchunks;
  int tile0;             // this launch covers the pair tiles [tile0, tile0 + t2) of the triangle (M > ~200: the per-thread
                         // accumulators of all tiles do not fit in shared memory, so the triangle is split over launches)
};

// Dynamic shared memory (bytes): acc[npass*TC*4] f64 | stage[kStages][chunk*(mp+QP)] f64 | zs[2*mt*QP] f64 |
//                                tiles[npass*TC] u32 | full[kStages], empty[kStages] u64 | etab[256] f64 (2 048 for EXPV 8)
template <int QP, int EXPV>
__global__ void __launch_bounds__(384, 1) psi2_fwd_kernel(Psi2FwdParams p) {
  extern __shared__ __align__(16) double sm[];
  const int T = blockDim.x, tid = threadIdx.x, TC = T - 32, ncw = TC / 32;
  double* acc = sm;
  double* stage = acc + (size_t)p.npass * TC * 4;
  const size_t stage_len = (size_t)p.chunk * (p.mp + QP);
  double* zs = stage + kStages * stage_len;
  unsigned* tiles = reinterpret_cast<unsigned*>(zs + 2 * p.mt * QP);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (((size_t)p.npass * TC + 1) & ~(size_t)1));
  uint64_t* empty = full + kStages;
  double* etab = reinterpret_cast<double*>(empty + kStages);

  for (int i = tid; i < 2 * p.mt * QP; i += T) {
    int m = i / QP, q = i % QP;
    zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0;
  }
  if (EXPV == 8) { for (int i = threadIdx.x; i < kExpTabSizeFwd; i += blockDim.x) etab[i] = p.exptab[i]; }
  else if (EXPV >= 4) load_exp_table(etab, p.exptab);
  for (int i = tid; i < p.npass * TC; i += T) {
    int ti, tj; tile_from_index(i < p.t2 ? i + p.tile0 : 0, p.mt, ti, tj);
    tiles[i] = (unsigned)(2 * ti) | ((unsigned)(2 * tj) << 16);
  }
  for (int i = tid; i < p.npass * TC * 4; i += T) acc[i] = 0.0;
  for (int i = tid; i < p.nseg; i += T) p.tags[blockIdx.x * p.nseg + i] = -1;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncw); }
    mbar_fence_init();
  }
  __syncthreads();
  const int64_t items = p.nchunks * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  if (lo >= hi) return;

  if (tid >= TC) {
    // ------------------------------------------------------------------ producer warp (one elected lane)
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
  // -------------------------------------------------------------------------------- consumer warps
  Exp<EXPV> ex; ex.init(etab);
  const int slot_len = p.npass * TC * 4;
  auto flush = [&](int seg, int b) {                      // thread-private: no barrier needed
    double* dst = p.part + ((size_t)blockIdx.x * p.nseg + seg) * slot_len;
    for (int pass = 0; pass < p.npass; ++pass) {
      double* a = acc + ((size_t)pass * TC + tid) * 4;
      double* d = dst + ((size_t)pass * TC + tid) * 4;
      reinterpret_cast<double2*>(d)[0] = reinterpret_cast<double2*>(a)[0];
      reinterpret_cast<double2*>(d)[1] = reinterpret_cast<double2*>(a)[1];
      a[0] = 0; a[1] = 0; a[2] = 0; a[3] = 0;
    }
    if (tid == 0) p.tags[blockIdx.x * p.nseg + seg] = b;
  };
  int cur_b = (int)(lo / p.nchunks), seg = 0;
  for (int64_t item = lo; item < hi; ++item) {
    const int k = (int)(item - lo), s = k % kStages;
    const int b = (int)(item / p.nchunks);
    const int64_t n0 = (item % p.nchunks) * p.chunk;
    const int nc = (int)min((int64_t)p.chunk, p.n - n0);
    if (b != cur_b) { flush(seg++, cur_b); cur_b = b; }
    mbar_wait(&full[s], (k / kStages) & 1);
    const double* rt = stage + s * stage_len;
    const double* vt = rt + (size_t)p.chunk * p.mp;
    for (int pass = 0; pass < p.npass; ++pass) {
      const unsigned tl = tiles[pass * TC + tid];
      const int m0 = tl & 0xffff, c0 = tl >> 16;
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
      double av[4] = {0, 0, 0, 0};
      int n = 0;
#if DPGP_FWD_ROWS2
      // two rows per iteration: 8 exponent chains and 8 exp chains in lockstep
      DPGP_FWD_UNROLL(DPGP_XP_FWD_UNROLL)
      for (; n + 1 < nc; n += 2) {
        double ev[8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const double2 ra = *reinterpret_cast<const double2*>(rt + (n + u) * p.mp + m0);
          const double2 rc = *reinterpret_cast<const double2*>(rt + (n + u) * p.mp + c0);
          double vq[QP];
#pragma unroll
          for (int q = 0; q < QP; q += 2) {
            const double2 t2 = *reinterpret_cast<const double2*>(vt + (n + u) * QP + q);
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
          ev[4 * u] = e00; ev[4 * u + 1] = e01; ev[4 * u + 2] = e10; ev[4 * u + 3] = e11;
        }
        double a8[8] = {av[0], av[1], av[2], av[3], 0, 0, 0, 0};
        exp_acc_k<EXPV, 8>(ex, ev, a8);
        av[0] = a8[0] + a8[4]; av[1] = a8[1] + a8[5]; av[2] = a8[2] + a8[6]; av[3] = a8[3] + a8[7];
      }
#endif
#pragma unroll 1
      for (; n < nc; ++n) {
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
        const double ev[4] = {e00, e01, e10, e11};
        exp_acc_k<EXPV, 4>(ex, ev, av);
      }
      double* a = acc + ((size_t)pass * TC + tid) * 4;
      a[0] += av[0]; a[1] += av[1]; a[2] += av[2]; a[3] += av[3];
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(&empty[s]);
  }
  flush(seg, cur_b);
}

// Deterministic reduction of the per-CTA partials (fixed slot order) into the symmetric Psi2 [B,M,M].
struct Psi2ReduceParams {
  const double* part; const int* tags; double* psi2;
  int grid, nseg, slot_len, m, mt, t2, b, tile0; int64_t nchunks;
};
static __global__ void psi2_reduce_kernel(Psi2ReduceParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // (b, tile, e)
  if (idx >= p.b * p.t2 * 4) return;
  const int b = idx / (p.t2 * 4), rem = idx % (p.t2 * 4), t = rem >> 2, e = rem & 3;
  int ti, tj; tile_from_index(t + p.tile0, p.mt, ti, tj);
  const int m = 2 * ti + (e >> 1), c = 2 * tj + (e & 1);
  if (m >= p.m || c >= p.m || m > c) return;
  int c_lo, c_hi; cta_range_of_cluster(b, p.nchunks, p.b, p.grid, c_lo, c_hi);      // psi1.cuh; only these CTAs can hold cluster b
  double s = 0;
  for (int k = c_lo * p.nseg; k < (c_hi + 1) * p.nseg; ++k)
    if (p.tags[k] == b) s += p.part[(size_t)k * p.slot_len + rem];
  p.psi2[((size_t)b * p.m + m) * p.m + c] = s;
  p.psi2[((size_t)b * p.m + c) * p.m + m] = s;
}

}  // namespace dpgp
