// Fused backward of the psi2 statistic with the two big reductions on the 5th-generation tensor cores (bwd_variant 7).
//
//   g_np  = Gs_p exp(r_nm + r_nm' + sum_q v_nq D_pq)        (psi2_bwd_fused.cuh; replaces tf.gradients through
//   dr_nm = sum_{m'} g_n(m,m')   dv_nq = sum_p g_np D_pq     src/kernels/rbf_kernel.py:164-199)
//   dD_pq = sum_n g_np v_nq
//
// psi2_bwd_fused_kernel spends 20 of its 46 FP64 issues per unit on dv and dD -- two genuine GEMMs (K = pairs, K = rows) whose
// left operand g is produced one element per unit.  tcgen05 has no FP64 kind, but it has an exact one: kind::i8 with int32
// accumulators in TMEM.  So g is written ONCE, as 48-bit fixed point cut into 6 byte planes, and both products run as
// slice-by-slice integer MMAs (an Ozaki-style split):
//
//   scaling    g_np = w_p S_n g'_np,  S_n = exp(2 max_m r_nm),  g' = exp((r_nm - rmax_n) + (r_nm' - rmax_n) + v.D) in (0, 1]
//              (1 is attained on the diagonal pair of the row's best m, so the 48 bits sit right below each row's largest term).
//              The row scale and the cotangent move into the SMALL operands:  dD_pq = w_p sum_n g'_np (S_n v_nq),
//              dv_nq = S_n sum_p g'_np (w_p D_pq);  dr keeps full FP64 (two DFMAs per unit with w_p as the multiplier).
//   slices     A = g': 6 unsigned byte planes of round(g' (2^48 - 2^16)); B = S v / max|S v| per 64-row item and
//              w D / max|w D| per cluster: 7 signed byte digits of a 54-bit fixed point (x + 0x80..80, bytes ^ 0x80).
//              Products A_i B_j with i + j < 7 are kept (the rest is below 2^-53 of the largest term); the B slices are
//              stacked along N so ONE MMA multiplies plane i with slices 0 .. 6 - i into accumulator columns 16 i ..:
//              12 + 12 MMAs (M 128, N 112 .. 32, K 32) per stage of 64 pairs x 64 rows.
//   one tile,  the byte tile g'[pair][row] is stored in [8 pairs][16 rows] core blocks; the SAME bytes are the K-major A
//   two views  operand of dD (M = pairs, K = rows) and the MN-major A operand of dv (M = rows, K = pairs)
//              (csrc/umma.cuh; pinned on the hardware by csrc/microbench/umma_i8_probe.cu).
//
// Roles (384 threads, setmaxnreg): warps 0-7 producers -- the first phase of psi2_bwd_fused_kernel without the dv FMAs, the g
// tile and the whole second phase: exponent, exp, two DFMAs for dr, one DFMA + 6 PRMT + 6 STS.U16 per pair of rows for the
// planes (22 FP64 issues per unit instead of 46); warp 10 issues the MMAs; warps 8-9 drain the dD accumulators of a stage
// (TMEM -> registers, Horner over the 7 levels, contraction with 2 (z_m - z_m') into the per-warp dz slices of bwd_variant 6)
// and, once per item, the dv accumulators.  Stages are double-buffered in shared memory and in TMEM; all hand-overs are
// mbarriers (producers -> MMA: `full`; tcgen05.commit -> producers and drain warps: `done`; drain -> MMA: `ddempty`, `dvempty`).
// Every reduction order is fixed, integer sums are exact: results are bitwise reproducible.
#pragma once
#include "psi2_bwd_fused.cuh"
#include "umma.cuh"

namespace dpgp {

constexpr int kUmNSA = 6, kUmNSB = 7, kUmLV = 7;
constexpr int kUmRows = 64, kUmStagePairs = 64;
constexpr int kUmSP = 128, kUmSR = kUmStagePairs / 8 * 128 + 16;      // pair-group / row-group strides of the core blocks
constexpr int kUmPlane = kUmRows / 16 * kUmSR;                        // one byte plane of a stage
constexpr int kUmGStage = kUmNSA * kUmPlane;
constexpr int kUmDPlane = kUmStagePairs * 16;                         // w D slices: [slice][pair][16 q] (MN-major B)
constexpr int kUmDStage = kUmNSB * kUmDPlane;
constexpr int kUmVLbo = 16 * kUmNSB / 8 * 128;                        // S v slices: [16 rows chunk][n = 16 j + q][16 rows] (K-major B)
constexpr int kUmVBytes = kUmRows / 16 * kUmVLbo;
constexpr int kUmThreads = 384, kUmTmemCols = 512;
constexpr int kUmAccDD = 128, kUmAccDV = 256;                         // TMEM columns: dD stage s at 128 s, dv at 256
constexpr int kUmRS = 66;                                             // row stride of the transposed r / dr tiles (doubles)
constexpr int kUmProducerRegs = 200, kUmOtherRegs = 104;   // (200 - 168) * 256 <= (168 - 104) * 128: the pool only holds what the other warps release

struct Psi2BwdUmmaParams {
  Psi2BwdFusedParams f;          // part: per-warp dz slices as bwd_variant 6 (slices 0, 1 of every CTA are used)
  const double* wtab;            // [B][nrounds * 512]   symmetrised cotangent per schedule slot (round, warp, i, k)
  const unsigned char* dprime;   // [B][nrounds * 8][NSB][64][16]   signed byte digits of w D / scale_d
  const double* scale_d;         // [B]
  long long* prof;               // development: per-role cycle counters of CTA 0 (nullptr: off)
  int dbg;                       // development switches (DPGP_UM_SKIP): 1 no MMAs, 2 no drain epilogue, 16 no column-side REDs, 32 no dD MMAs, 64 no dv MMAs
};

__host__ __device__ inline size_t umma_smem_bytes(int mp, int qp) {
  return 2 * (size_t)kUmGStage + 2 * (size_t)kUmDStage + kUmVBytes +
         (2 * (size_t)mp * kUmRS + (size_t)mp * qp + kExpTabSize + 8 * 8 * (qp + 2) + 64 + 128 + 16) * 8 + 8 * 8 + 16;
}

template <int N> __device__ __forceinline__ void um_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void um_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void um_producer_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned um_prmt(unsigned a, unsigned b, unsigned sel) {
  unsigned r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r;
}
__device__ __forceinline__ void um_sts16(unsigned addr, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory"); }

__host__ __device__ inline double um_ca() { return 281474976710656.0 - 65536.0; }                 // 2^48 - 2^16
__host__ __device__ inline double um_cb() { return 18014398509481984.0; }                            // 2^54 = 2^(8 NSB - 2)
__host__ __device__ inline double um_kd() { return 6.103515625e-05 / (1.0 - 2.3283064365386963e-10); }  // 256^(NSA+NSB-2) / (CA CB) = 2^88 / (2^48 (1 - 2^-32) 2^54)
constexpr long long kUmOff = 0x80808080808080LL;                                                      // 128 in each of the NSB bytes

// NSB signed byte digits (most significant first) of round(x 2^54), |x| <= 1
__device__ __forceinline__ unsigned long long um_digits(double x) {
  const long long X = __double2ll_rn(x * um_cb());
  return (unsigned long long)(X + kUmOff) ^ (unsigned long long)kUmOff;
}

// out[k] = exp(x[k]), K chains in lockstep (the table variant without the weight multiply)
template <int EXPV, int K>
__device__ __forceinline__ void um_exp_k(const Exp<EXPV>& ex, const double (&x)[K], double (&out)[K]) {
  if constexpr (EXPV >= 4) {
    constexpr int BITS = exp_tab_bits(EXPV);
    using C = ExpTabConst<BITS>;
    double t[K], r[K], q[K], T[K];
#pragma unroll
    for (int k = 0; k < K; ++k) t[k] = fma(x[k], C::L2E_S, C::MAGIC);
#pragma unroll
    for (int k = 0; k < K; ++k) { T[k] = exp_tab_entry<BITS>(ex.stab, t[k]); t[k] -= C::MAGIC; }
#pragma unroll
    for (int k = 0; k < K; ++k) r[k] = fma(t[k], C::NLN2_S, x[k]);
    exp_tab_poly<BITS, K>(r, q);
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = fma(T[k] * r[k], q[k], T[k]);
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = ex.value(x[k]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Per-cluster tables: cotangent per schedule slot, scale of w D, byte digits of w D / scale in the stage layout.
struct UmmaTablesParams {
  const double* gbar; const double* z; const unsigned short* sched;
  double* wtab; unsigned char* dprime; double* scale_d;
  int q, m, nrounds;
};
static __global__ void __launch_bounds__(256) umma_tables_kernel(UmmaTablesParams p) {
  __shared__ double red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double* gb = p.gbar + (size_t)b * p.m * p.m;
  const int nslots = p.nrounds * 512;
  double mx = 0.0;
  for (int slot = tid; slot < nslots; slot += 256) {
    const int round = slot >> 9, w8 = (slot >> 6) & 7, i = (slot >> 3) & 7, k = slot & 7;
    const unsigned short it = p.sched[round * 8 + w8];
    double w = 0.0;
    if (it != kSchedIdle) {
      const int m = 8 * (it >> 8) + i, c = 8 * (it & 255) + k;
      w = sym_cotangent(gb, m, c, p.m);
      if (w != 0.0)
        for (int q = 0; q < p.q; ++q) { const double d = p.z[m * p.q + q] - p.z[c * p.q + q]; mx = fmax(mx, fabs(w) * d * d); }
    }
    p.wtab[(size_t)b * nslots + slot] = w;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = 0.0;
  for (int i = 0; i < 8; ++i) mx = fmax(mx, red[i]);
  const double scale = mx > 0.0 ? mx : 1.0;
  if (tid == 0) p.scale_d[b] = scale;
  const double inv = 1.0 / scale;
  for (int slot = tid; slot < nslots; slot += 256) {
    const int round = slot >> 9, w8 = (slot >> 6) & 7, i = (slot >> 3) & 7, k = slot & 7;
    const unsigned short it = p.sched[round * 8 + w8];
    unsigned long long dg[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) dg[q] = 0ull;
    if (it != kSchedIdle) {
      const int m = 8 * (it >> 8) + i, c = 8 * (it & 255) + k;
      const double w = sym_cotangent(gb, m, c, p.m);
      if (w != 0.0) {
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (q < p.q) { const double d = p.z[m * p.q + q] - p.z[c * p.q + q]; dg[q] = um_digits(w * d * d * inv); }
      }
    }
    unsigned char* dst = p.dprime + (((size_t)b * p.nrounds * 8 + (size_t)round * 8 + i) * kUmNSB) * kUmDPlane + (w8 * 8 + k) * 16;
#pragma unroll
    for (int j = 0; j < kUmNSB; ++j) {
      unsigned wd[4];
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        unsigned v = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) v |= (unsigned)((dg[4 * g4 + e] >> (8 * (kUmNSB - 1 - j))) & 0xffull) << (8 * e);
        wd[g4] = v;
      }
      *reinterpret_cast<uint4*>(dst + (size_t)j * kUmDPlane) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct UmSmem {
  unsigned char* gst; unsigned char* dst; unsigned char* vst;
  double* rT; double* drT; double* zs; double* etab; double* dtab; double* rhat; double* sn; double* scal;
  uint64_t* full; uint64_t* done; uint64_t* ddempty; uint64_t* dvempty;
  uint32_t* tmem_slot;
};

// Horner over the accumulator levels of `cols0`: h[q] = sum_l acc_l[q] 256^-l
template <int QP>
__device__ __forceinline__ void um_read_levels(uint32_t taddr, double (&h)[QP]) {
#pragma unroll
  for (int q = 0; q < QP; ++q) h[q] = 0.0;
#pragma unroll
  for (int l = kUmLV - 1; l >= 0; --l) {
    uint32_t r[16];
    tmem_ld16(taddr + 16 * l, r);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < QP; ++q) h[q] = fma(h[q], 0.00390625, (double)(int)r[q]);
  }
}

template <int QP, int EXPV>
__device__ __forceinline__ void um_producer(const Psi2BwdUmmaParams& pp, const UmSmem& S) {
  const Psi2BwdFusedParams& p = pp.f;
  constexpr int RS = kUmRS, DS = QP + 2, QH = QP / 2, ROWS = kUmRows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Exp<EXPV> ex; ex.init(S.etab);
  double* dtw = S.dtab + (size_t)warp * 8 * DS;
  const int64_t items = p.ngroups * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  uint32_t T = 0; int par = 0;
  const double CA = um_ca(), MAGIC52 = 4503599627370496.0;
  long long prof_wait = 0; const long long prof_t0 = clock64();
  for (int64_t item = lo; item < hi; ++item, par ^= 1) {
    const int b = (int)(item / p.ngroups);
    const int64_t n0 = (item % p.ngroups) * ROWS;
    const int nc = (int)min((int64_t)ROWS, p.n - n0);
    // every MMA of the previous item has completed (they read the S v slices that are rebuilt below)
    if (T >= 1) mbar_wait(&S.done[(T - 1) & 1], ((T - 1) >> 1) & 1);
    if (T >= 2) mbar_wait(&S.done[(T - 2) & 1], ((T - 2) >> 1) & 1);
    um_producer_barrier();                               // the previous item's dr drain has finished with rT / drT
    {
      const double* src = p.r + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < ROWS * p.mp; idx += 256) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        S.rT[(size_t)m * RS + row] = (row < nc) ? __ldcs(src + idx) : kRClamp;
      }
      for (int idx = tid; idx < p.mp * RS; idx += 256) S.drT[idx] = 0.0;
    }
    um_producer_barrier();
    {                                                    // row maximum of r, row scale S_n = exp(2 rmax)
      const int row = tid >> 2, part = tid & 3;
      double mx = -1.0e300;
      for (int m = part; m < p.m; m += 4) mx = fmax(mx, S.rT[(size_t)m * RS + row]);      // real inducing points only
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      if (part == 0) { S.rhat[row] = mx; S.sn[par * 64 + row] = (row < nc) ? exp(2.0 * mx) : 0.0; }
    }
    um_producer_barrier();
    for (int idx = tid; idx < p.mp * ROWS; idx += 256) { const int m = idx >> 6, row = idx & 63; S.rT[(size_t)m * RS + row] -= S.rhat[row]; }
    // S v slices of the item (K-major B operand of the dD products)
    const double* vsrc = p.v + ((int64_t)b * p.n + n0) * QP;
    {
      double mx = 0.0;
      for (int idx = tid; idx < ROWS * QP; idx += 256) { const int row = idx / QP; if (row < nc) mx = fmax(mx, fabs(vsrc[idx]) * S.sn[par * 64 + row]); }
#pragma unroll
      for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) S.scal[4 + warp] = mx;
    }
    um_producer_barrier();
    {
      double mx = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) mx = fmax(mx, S.scal[4 + w]);
      const double scale_v = mx > 0.0 ? mx : 1.0, inv = 1.0 / scale_v;
      if (tid == 0) S.scal[par] = scale_v;
      for (int idx = tid; idx < ROWS * QP; idx += 256) {
        const int row = idx / QP, q = idx - row * QP;
        const double x = (row < nc) ? vsrc[idx] * S.sn[par * 64 + row] * inv : 0.0;
        const unsigned long long dg = um_digits(x);
        unsigned char* dst = S.vst + (row >> 4) * kUmVLbo + (row & 15);
#pragma unroll
        for (int j = 0; j < kUmNSB; ++j) { const int n = 16 * j + q; dst[(n >> 3) * 128 + (n & 7) * 16] = (unsigned char)(dg >> (8 * (kUmNSB - 1 - j))); }
      }
    }
    fence_proxy_async_smem();
    double vq[2][QP];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int row = 2 * lane + rr;
      const double* vs = vsrc + (size_t)(row < nc ? row : 0) * QP;
#pragma unroll
      for (int q = 0; q < QP; q += 2) {
        const double2 t2 = __ldcs(reinterpret_cast<const double2*>(vs + q));
        vq[rr][q] = (row < nc) ? t2.x : 0.0; vq[rr][q + 1] = (row < nc) ? t2.y : 0.0;
      }
    }
    um_producer_barrier();
    double wn0, wn1;
    {
      const double* w0 = pp.wtab + (size_t)b * p.nrounds * 512 + warp * 64;
      wn0 = __ldg(w0 + lane); wn1 = __ldg(w0 + 32 + lane);
    }
    uint4 dnext[2];                                      // byte digits of w D for the NEXT stage (NSB x 128 bytes per warp, from L2)
    auto load_digits = [&](const unsigned char* src) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = lane + 32 * e;
        if (idx < kUmNSB * 8) dnext[e] = __ldg(reinterpret_cast<const uint4*>(src + (idx >> 3) * kUmDPlane + (idx & 7) * 16));
      }
    };
    load_digits(pp.dprime + (size_t)b * p.nrounds * 8 * kUmDStage + warp * 128);

    for (int round = 0; round < p.nrounds; ++round) {
      const unsigned short it = p.sched[round * kFusedWarps + warp];
      const int bi = it >> 8, bj = it & 255;
      const double* wrow = pp.wtab + ((size_t)b * p.nrounds + round) * 512 + warp * 64;
      const unsigned char* dsrc = pp.dprime + ((size_t)b * p.nrounds + round) * 8 * kUmDStage + warp * 128;
      // cotangents of the block's 64 pairs, two per lane; those of the next round travel while this one computes
      const double wc0 = wn0, wc1 = wn1;
      if (round + 1 < p.nrounds) { wn0 = __ldg(wrow + 512 + lane); wn1 = __ldg(wrow + 512 + 32 + lane); }
      double cs[8][2];
#pragma unroll
      for (int k = 0; k < 8; ++k) { cs[k][0] = 0.0; cs[k][1] = 0.0; }
#pragma unroll 1
      for (int i = 0; i < 8; ++i, ++T) {
        const uint32_t s = T & 1, u = T >> 1;
        const uint4 dld[2] = {dnext[0], dnext[1]};
        if (i < 7 || round + 1 < p.nrounds) load_digits(dsrc + (size_t)(i + 1) * kUmDStage);      // (i = 7: first stage of the next round)
        const double wsel = __shfl_sync(0xffffffffu, (i & 4) ? wc1 : wc0, 8 * (i & 3) + (lane >> 1));
        const long long tw0 = clock64();
        if (u >= 1) mbar_wait(&S.done[s], (u - 1) & 1);          // the MMAs that read this stage's previous contents are done
        prof_wait += clock64() - tw0;
        unsigned char* gs = S.gst + s * kUmGStage;
        if (it != kSchedIdle) {
          const int m = 8 * bi + i;
          if (lane < 16) {
            const int k = lane >> 1, qh = lane & 1, c = 8 * bj + k;
#pragma unroll
            for (int j = 0; j < QH; ++j) {
              const int q = qh * QH + j;
              const double d = S.zs[m * QP + q] - S.zs[c * QP + q];
              dtw[k * DS + q] = d * d;
            }
            if (qh == 0) { dtw[k * DS + QP] = wsel; dtw[k * DS + QP + 1] = 0.0; }
          }
          __syncwarp();
          const double2 rm = *reinterpret_cast<const double2*>(S.rT + (size_t)m * RS + 2 * lane);
          double rs0 = 0.0, rs1 = 0.0;
          const unsigned gaddr = smem_u32(gs) + (lane >> 3) * kUmSR + warp * kUmSP + 2 * (lane & 7);
#pragma unroll
          for (int k0 = 0; k0 < 8; k0 += 4) {
            double e[8], g[8], wv[4];
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              const double* dt = dtw + (k0 + u4) * DS;
              double dq[QP];
#pragma unroll
              for (int q = 0; q < QP; q += 2) { const double2 t2 = *reinterpret_cast<const double2*>(dt + q); dq[q] = t2.x; dq[q + 1] = t2.y; }
              wv[u4] = dt[QP];
              const double2 rc = *reinterpret_cast<const double2*>(S.rT + (size_t)(8 * bj + k0 + u4) * RS + 2 * lane);
              double ea0 = rm.x, eb0 = rc.x, ea1 = rm.y, eb1 = rc.y;
#pragma unroll
              for (int q = 0; q < QP; q += 2) {
                ea0 = fma(vq[0][q], dq[q], ea0); eb0 = fma(vq[0][q + 1], dq[q + 1], eb0);
                ea1 = fma(vq[1][q], dq[q], ea1); eb1 = fma(vq[1][q + 1], dq[q + 1], eb1);
              }
              e[2 * u4] = ea0 + eb0; e[2 * u4 + 1] = ea1 + eb1;
            }
            um_exp_k<EXPV, 8>(ex, e, g);
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              const int k = k0 + u4;
              rs0 = fma(wv[u4], g[2 * u4], rs0); rs1 = fma(wv[u4], g[2 * u4 + 1], rs1);
              cs[k][0] = fma(wv[u4], g[2 * u4], cs[k][0]); cs[k][1] = fma(wv[u4], g[2 * u4 + 1], cs[k][1]);
              const double t0 = fma(g[2 * u4], CA, MAGIC52), t1 = fma(g[2 * u4 + 1], CA, MAGIC52);
              const unsigned lo0 = (unsigned)__double2loint(t0), hi0 = (unsigned)__double2hiint(t0);
              const unsigned lo1 = (unsigned)__double2loint(t1), hi1 = (unsigned)__double2hiint(t1);
              const unsigned a = gaddr + k * 16;
              um_sts16(a + 0 * kUmPlane, um_prmt(hi0, hi1, 0x51));      // plane 0 = most significant byte (bits 40..47)
              um_sts16(a + 1 * kUmPlane, um_prmt(hi0, hi1, 0x40));
              um_sts16(a + 2 * kUmPlane, um_prmt(lo0, lo1, 0x73));
              um_sts16(a + 3 * kUmPlane, um_prmt(lo0, lo1, 0x62));
              um_sts16(a + 4 * kUmPlane, um_prmt(lo0, lo1, 0x51));
              um_sts16(a + 5 * kUmPlane, um_prmt(lo0, lo1, 0x40));
            }
          }
          double2* dr = reinterpret_cast<double2*>(S.drT + (size_t)m * RS + 2 * lane);
          double2 acc = *dr; acc.x += rs0; acc.y += rs1; *dr = acc;
        }
        {
          unsigned char* dd = S.dst + s * kUmDStage + warp * 128;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int idx = lane + 32 * e;
            if (idx < kUmNSB * 8) *reinterpret_cast<uint4*>(dd + (idx >> 3) * kUmDPlane + (idx & 7) * 16) = dld[e];
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.full[s]);
      }
      if (it != kSchedIdle) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          double2* dc = reinterpret_cast<double2*>(S.drT + (size_t)(8 * bj + k) * RS + 2 * lane);
          double2 acc = *dc; acc.x += cs[k][0]; acc.y += cs[k][1]; *dc = acc;
        }
      }
      um_producer_barrier();
    }
    {
      double* dst = p.dr + ((int64_t)b * p.n + n0) * p.mp;
      for (int idx = tid; idx < nc * p.mp; idx += 256) {
        const int row = idx / p.mp, m = idx - row * p.mp;
        __stcs(dst + idx, S.drT[(size_t)m * RS + row] * S.sn[par * 64 + row]);
      }
    }
  }
  if (pp.prof && blockIdx.x == 0 && lane == 0) { pp.prof[2 * warp] = clock64() - prof_t0; pp.prof[2 * warp + 1] = prof_wait; }
}

template <int QP>
__device__ __forceinline__ void um_mma(const Psi2BwdUmmaParams& pp, const UmSmem& S, uint32_t tm) {
  const Psi2BwdFusedParams& p = pp.f;
  const int64_t items = p.ngroups * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  const int nst = p.nrounds * 8;
  uint64_t dd_a[2], dv_a[2], dv_b[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    dd_a[s] = umma_smem_desc(S.gst + s * kUmGStage, kUmSR, kUmSP);        // K-major view of the byte tile: M = pairs, K = rows
    dv_a[s] = umma_smem_desc(S.gst + s * kUmGStage, kUmSP, kUmSR);        // MN-major view: M = rows, K = pairs
    dv_b[s] = umma_smem_desc(S.dst + s * kUmDStage, 128, kUmDPlane);
  }
  const uint64_t dd_b = umma_smem_desc(S.vst, kUmVLbo, 128);
  uint32_t T = 0, item_idx = 0;
  long long pf_full = 0, pf_dd = 0; const long long pf_t0 = clock64();
  for (int64_t item = lo; item < hi; ++item, ++item_idx) {
    for (int st = 0; st < nst; ++st, ++T) {
      const uint32_t s = T & 1, u = T >> 1;
      long long tq = clock64();
      mbar_wait(&S.full[s], u & 1);
      pf_full += clock64() - tq; tq = clock64();
      if (u >= 1) mbar_wait(&S.ddempty[s], (u - 1) & 1);
      pf_dd += clock64() - tq;
      if (st == 0 && item_idx >= 1) mbar_wait(&S.dvempty[0], (item_idx - 1) & 1);
      tc_fence_after();
      if (pp.dbg & 1) { umma_commit(&S.done[s]); continue; }
      // descriptors differ from the stage's base descriptors by constant address offsets (16-byte units in the low bits)
      const uint64_t a_dd = dd_a[s], a_dv = dv_a[s], b_dv = dv_b[s];
      const uint32_t acc_dd = tm + kUmAccDD * s;
      if (!(pp.dbg & 32)) {
#pragma unroll
        for (int ks = 0; ks < kUmRows / 32; ++ks)
#pragma unroll
          for (int i = 0; i < kUmNSA; ++i) {
            constexpr int NJ0 = kUmLV < kUmNSB ? kUmLV : kUmNSB;
            const int nj = (kUmLV - i) < NJ0 ? (kUmLV - i) : NJ0;
            umma_i8(acc_dd + 16 * i, a_dd + (uint64_t)((i * kUmPlane + ks * 2 * kUmSR) >> 4), dd_b + (uint64_t)((ks * 2 * kUmVLbo) >> 4),
                    umma_idesc_i8(128, 16 * nj, false, true, false, false), !(ks == 0 && i == 0));
          }
      }
      if (!(pp.dbg & 64)) {
        const bool first = st == 0;
#pragma unroll
        for (int ks = 0; ks < kUmStagePairs / 32; ++ks)
#pragma unroll
          for (int i = 0; i < kUmNSA; ++i) {
            constexpr int NJ0 = kUmLV < kUmNSB ? kUmLV : kUmNSB;
            const int nj = (kUmLV - i) < NJ0 ? (kUmLV - i) : NJ0;
            umma_i8(tm + kUmAccDV + 16 * i, a_dv + (uint64_t)((i * kUmPlane + ks * 4 * kUmSP) >> 4), b_dv + (uint64_t)((ks * 4 * 128) >> 4),
                    umma_idesc_i8(128, 16 * nj, false, true, true, true), !(first && ks == 0 && i == 0));
          }
      }
      umma_commit(&S.done[s]);
    }
  }
  if (pp.prof && blockIdx.x == 0) { pp.prof[16] = clock64() - pf_t0; pp.prof[17] = pf_full; pp.prof[18] = pf_dd; }
}

template <int QP>
__device__ __forceinline__ void um_drain(const Psi2BwdUmmaParams& pp, const UmSmem& S, uint32_t tm, int dw) {
  const Psi2BwdFusedParams& p = pp.f;
  const int lane = threadIdx.x & 31, wslot = 4 * dw + (lane >> 3), k = lane & 7;
  const int64_t items = p.ngroups * p.b;
  const int64_t lo = items * blockIdx.x / gridDim.x, hi = items * (blockIdx.x + 1) / gridDim.x;
  double* dzr = p.part + ((size_t)blockIdx.x * kFusedWarps + dw) * 2 * p.mp * QP;
  double* dzc = dzr + (size_t)p.mp * QP;
  const uint64_t keep = l2_evict_last_policy();
  const uint32_t lane_base = (uint32_t)(32 * dw) << 16;
  const double KD = um_kd();
  uint32_t T = 0; int par = 0;
  long long pd_wait = 0, pd_read = 0; const long long pd_t0 = clock64();
  for (int64_t item = lo; item < hi; ++item, par ^= 1) {
    const int b = (int)(item / p.ngroups);
    const int64_t n0 = (item % p.ngroups) * kUmRows;
    const int nc = (int)min((int64_t)kUmRows, p.n - n0);
    for (int round = 0; round < p.nrounds; ++round) {
      const unsigned short it = p.sched[round * kFusedWarps + wslot];
      const bool live = it != kSchedIdle;
      const int bi = live ? (it >> 8) : 0, bj = live ? (it & 255) : 0;
      const double* wrow = pp.wtab + ((size_t)b * p.nrounds + round) * 512 + wslot * 64 + k;
#pragma unroll 1
      for (int i = 0; i < 8; ++i, ++T) {
        const uint32_t s = T & 1, u = T >> 1;
        const double w = __ldg(wrow + i * 8);
        long long tq = clock64();
        mbar_wait(&S.done[s], u & 1);
        pd_wait += clock64() - tq; tq = clock64();
        tc_fence_after();
        double h[QP];
        um_read_levels<QP>(tm + lane_base + kUmAccDD * s, h);
        pd_read += clock64() - tq;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.ddempty[s]);
        if (pp.dbg & 2) continue;
        const double fac = w * S.scal[par] * KD;
        const int m = 8 * bi + i, c = 8 * bj + k;
        double t[QP], rsum[QP];
#pragma unroll
        for (int q = 0; q < QP; ++q) t[q] = (S.zs[m * QP + q] - S.zs[c * QP + q]) * (h[q] * fac);
#pragma unroll
        for (int q = 0; q < QP; ++q) {
          double a = t[q] + __shfl_xor_sync(0xffffffffu, t[q], 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          rsum[q] = a + __shfl_xor_sync(0xffffffffu, a, 4);
        }
        if (live && k == 0) {
#pragma unroll
          for (int q = 0; q < QP; ++q) red_add_f64_keep(dzr + (size_t)m * QP + q, rsum[q], keep);
        }
        if (w != 0.0 && !(pp.dbg & 16)) {
#pragma unroll
          for (int q = 0; q < QP; ++q) red_add_f64_keep(dzc + (size_t)c * QP + q, t[q], keep);
        }
      }
    }
    // dv of the item: the commit of its last stage covers every MMA into the dv accumulators
    {
      double h[QP];
      um_read_levels<QP>(tm + lane_base + kUmAccDV, h);
      const int row = 32 * dw + lane;
      const double fac = S.sn[par * 64 + row] * __ldg(pp.scale_d + b) * KD;
      if (row < nc) {
        double* dst = p.dv + ((int64_t)b * p.n + n0 + row) * QP;
#pragma unroll
        for (int q = 0; q < QP; ++q) __stcs(dst + q, h[q] * fac);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.dvempty[0]);
    }
  }
  if (pp.prof && blockIdx.x == 0 && lane == 0) { pp.prof[20 + 3 * dw] = clock64() - pd_t0; pp.prof[21 + 3 * dw] = pd_wait; pp.prof[22 + 3 * dw] = pd_read; }
}

template <int QP, int EXPV>
__global__ void __launch_bounds__(kUmThreads, 1) psi2_bwd_umma_kernel(Psi2BwdUmmaParams pp) {
  extern __shared__ __align__(128) unsigned char um_sm[];
  const Psi2BwdFusedParams& p = pp.f;
  UmSmem S;
  S.gst = um_sm; S.dst = S.gst + 2 * kUmGStage; S.vst = S.dst + 2 * kUmDStage;
  S.rT = reinterpret_cast<double*>(S.vst + kUmVBytes);
  S.drT = S.rT + (size_t)p.mp * kUmRS; S.zs = S.drT + (size_t)p.mp * kUmRS; S.etab = S.zs + (size_t)p.mp * QP;
  S.dtab = S.etab + kExpTabSize; S.rhat = S.dtab + 8 * 8 * (QP + 2); S.sn = S.rhat + 64; S.scal = S.sn + 128;
  S.full = reinterpret_cast<uint64_t*>(S.scal + 16); S.done = S.full + 2; S.ddempty = S.done + 2; S.dvempty = S.ddempty + 2;
  S.tmem_slot = reinterpret_cast<uint32_t*>(S.dvempty + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < p.mp * QP; i += kUmThreads) { const int m = i / QP, q = i % QP; S.zs[i] = (m < p.m && q < p.q) ? p.z[m * p.q + q] : 0.0; }
  for (int i = tid; i < kExpTabSize; i += kUmThreads) S.etab[i] = p.exptab[i];
  for (int i = tid; i < (2 * kUmGStage + 2 * kUmDStage + kUmVBytes) / 16; i += kUmThreads) reinterpret_cast<uint4*>(um_sm)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&S.full[s], 8); mbar_init(&S.done[s], 1); mbar_init(&S.ddempty[s], 2); }
    mbar_init(&S.dvempty[0], 2);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(S.tmem_slot, kUmTmemCols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *S.tmem_slot;
  if (warp < 8) {
    um_setmaxnreg_inc<kUmProducerRegs>();
    um_producer<QP, EXPV>(pp, S);
  } else {
    um_setmaxnreg_dec<kUmOtherRegs>();
    if (warp == 10) { if ((tid & 31) == 0) um_mma<QP>(pp, S, tm); __syncwarp(); }
    else if (warp < 10) um_drain<QP>(pp, S, tm, warp - 8);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_free(tm, kUmTmemCols);
}

}  // namespace dpgp
