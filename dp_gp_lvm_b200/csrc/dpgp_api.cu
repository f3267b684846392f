// C ABI of the DP-GP-LVM hot path (see include/dpgp.h).  Host-side orchestration only: every number is
// produced by the kernels in psi1.cuh / psi2.cuh / psi2_bwd_fused.cuh / chain2.cuh / bound.cuh / small.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dpgp.h"
#include "bound.cuh"
#include "small.cuh"
#include "common.cuh"
#include "qp_kernels.cuh"

using namespace dpgp;

namespace {
constexpr int kNumPhases = 9;
const char* kPhaseNames[kNumPhases] = {"prep", "psi2_fwd", "psi1_fwd", "bound", "psi2_bwd_n", "psi2_bwd_pair",
                                       "chain_bwd", "reduce", "psi2_bwd_fused"};
enum Phase { PH_PREP, PH_PSI2F, PH_PSI1F, PH_BOUND, PH_BWDN, PH_BWDP, PH_CHAIN, PH_REDUCE, PH_BWDF };
}  // namespace

struct dpgp_handle {
  int device = 0, sms = 0;
  int64_t n = 0;
  int d = 0, q = 0, qp = 0, m = 0, mp = 0, mt = 0, t2 = 0, b = 0, mode = 0, ncols = 0, cpad = 0;
  int expv = 2, grid = 0;
  const QpLaunchers* k = nullptr;
  // psi2 forward
  int f_threads = 0, f_npass = 0, f_chunk = 32, f_splits = 1, f_t2 = 0, f_nseg = 2, p1_nseg = 2, p1_grid = 0; size_t f_smem = 0;
  // psi2 backward (pair side)
  int p_threads = 0, p_jb = 0, p_ng = 0, p_chunk = 64, p_nseg = 2; size_t p_smem = 0;
  // psi2 backward (n side)
  int n_threads = 0; size_t n_smem = 0;
  // psi2 backward (fused): rows per lane, rounds of the block schedule, grid, cluster segments per CTA
  int chain_variant = 1, c2_rows = 32, c2_grid = 0, c2_groups = 1; size_t c2_smem = 0;
  double *c2_mu_part = nullptr, *c2_s_part = nullptr;
  int bwd_variant = 1, u_rows = 2, u_nrounds = 0, u_grid = 0, u_nseg = 1; size_t u_smem = 0, u_slice = 0;
  unsigned short* u_sched = nullptr; double* u_part = nullptr; int* u_tags = nullptr; double* exptab = nullptr;
  double* exptab_fwd = nullptr; int expv_fwd = 4;      // forward kernel: 2 048-entry table + degree-3 polynomial when the default variant is selected
  // bwd_variant 7 (tcgen05 int8 slice products): per-cluster cotangent / w D digit tables and their scale
  double* um_wtab = nullptr; unsigned char* um_dprime = nullptr; double* um_scale = nullptr; size_t um_smem = 0;
  // workspace
  std::vector<void*> allocs;
  size_t ws_bytes = 0;
  double *r = nullptr, *v = nullptr, *bco = nullptr, *dv = nullptr;
  double *f_part = nullptr, *p1_part = nullptr, *cs_part = nullptr, *bp_part = nullptr, *ddsym = nullptr;
  int *f_tags = nullptr, *p1_tags = nullptr, *bp_tags = nullptr, *bad = nullptr;
  double *zpart = nullptr;       // [B][ceil(M / 8)][Q + 1] partials of zchain_kernel
  double *fb = nullptr, *dk = nullptr, *dzk = nullptr, *dzd = nullptr, *dadirect = nullptr;
  // M x M chain (bound.cuh): inverse factors, dense intermediates [B][M][M] / [B][M][C], scalars of the second factorisation
  double *lk = nullptr, *la = nullptr, *c1 = nullptr;      // the two Cholesky factors [B][M][M]; L^-1 P [B][M][C]
  double *linv = nullptr, *kinv = nullptr, *t1 = nullptr, *hmat = nullptr, *lainv = nullptr, *rmat = nullptr, *smat = nullptr;
  double *gmat = nullptr, *rtg = nullptr, *gtg = nullptr, *wmat = nullptr, *cm = nullptr, *umat = nullptr, *pu = nullptr, *fscal = nullptr;
  double *fwork = nullptr, *ftmp = nullptr;          // global-memory factorisation (M > kFacMaxM)
  // K_uu factor on a side stream: started by dpgp_stats_fwd, joined by dpgp_bound
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_kuu = nullptr;
  bool force_global_factor = false;
  struct Guarded { char* base; size_t bytes, rounded; };
  bool guard = false; std::vector<Guarded> guarded;
  bool fine = false; cudaStream_t fine_stream = nullptr; std::vector<std::pair<const char*, cudaEvent_t>> fine_ev;
  bool kuu_pending = false; const double *kuu_z = nullptr, *kuu_gamma = nullptr, *kuu_alpha = nullptr;
  double *dzp = nullptr, *dgp = nullptr, *dap = nullptr, *dummy = nullptr, *dtab = nullptr, *gtab = nullptr;
  int cs_grid = 0;
  int64_t launches = 0;
  std::string err;
  // timing
  bool timing = false;
  cudaEvent_t ev0[kNumPhases] = {}, ev1[kNumPhases] = {};
  bool ev_used[kNumPhases] = {};
};

namespace {

int fail(dpgp_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (h) h->err = buf;
  return code;
}
#define CU(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
  return fail(h, DPGP_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)
#define POST_LAUNCH(h, name) do { ++(h)->launches; cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) \
  return fail(h, DPGP_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
  if ((h)->fine) fine_mark(h, name); } while (0)

// Makes the handle's device current for the duration of a call and restores the caller's device afterwards (a handle may
// be created, used or destroyed -- e.g. from a garbage collector -- while another device is current).
struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// Development aid (dpgp_debug_launch_times): one event after every launch, on the stream of the last hot-path call.
void fine_mark(dpgp_handle* h, const char* name);

// Workspace allocation.  With DPGP_GUARD set in the environment at dpgp_create time every buffer is bracketed by two
// kGuardBytes bands of a fixed byte pattern, which dpgp_check_guards verifies: an out-of-bounds WRITE of any kernel into the
// neighbourhood of a workspace buffer is caught (compute-sanitizer is not available on every pool; tests/test_gpu_guards.py).
constexpr size_t kGuardBytes = 4096;
constexpr int kGuardPattern = 0xA5;
template <typename T>
int ws_alloc(dpgp_handle* h, T** p, size_t count) {
  void* ptr = nullptr;
  size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
  const size_t pad = h->guard ? kGuardBytes : 0;
  const size_t rounded = (bytes + 255) / 256 * 256;                 // the upper band starts right after the (rounded) buffer
  cudaError_t e = cudaMalloc(&ptr, rounded + 2 * pad);
  if (e != cudaSuccess) return fail(h, DPGP_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  h->allocs.push_back(ptr); h->ws_bytes += bytes; *p = (T*)((char*)ptr + pad);
  if (h->guard) {
    cudaMemset(ptr, kGuardPattern, pad);
    cudaMemset((char*)ptr + pad + bytes, kGuardPattern, rounded - bytes + pad);
    h->guarded.push_back({(char*)ptr, bytes, rounded});
  }
  return DPGP_OK;
}

int pad_q(int q) { return q <= 12 ? round_up(q, 2) : round_up(q, 4); }      // instantiated: 2, 4, ..., 12, 16, 20, ..., 32

struct PhaseTimer {
  dpgp_handle* h; int ph; cudaStream_t st;
  PhaseTimer(dpgp_handle* h_, int ph_, cudaStream_t st_) : h(h_), ph(ph_), st(st_) {
    if (h->timing) { cudaEventRecord(h->ev0[ph], st); }
  }
  ~PhaseTimer() { if (h->timing) { cudaEventRecord(h->ev1[ph], st); h->ev_used[ph] = true; } }
};

const QpLaunchers* launchers_for(int qp) {
  switch (qp) {
#define DPGP_QP_CASE(q) case q: return qp_launchers_##q();
    DPGP_QP_LIST(DPGP_QP_CASE)
#undef DPGP_QP_CASE
    default: return nullptr;
  }
}
#ifdef DPGP_EXPERIMENTAL
constexpr bool kExperimental = true;
#else
constexpr bool kExperimental = false;
#endif


// Rounds of 8x8 pair blocks for psi2_bwd_fused_kernel: round-robin tournament on the nb m-blocks (circle method).
// Every round holds <= kFusedWarps blocks that touch pairwise disjoint m-blocks; with nb odd the m-block that
// sits out a round does its diagonal block there, otherwise the diagonal blocks fill extra rounds.
std::vector<unsigned short> build_fused_schedule(int nb, int* nrounds) {
  std::vector<std::vector<unsigned short>> rounds;
  const int nbe = nb + (nb & 1);
  std::vector<char> diag_done(nb, 0);
  for (int r = 0; r + 1 < nbe; ++r) {
    std::vector<unsigned short> items;
    auto add = [&](int a, int c) {
      if (a >= nb || c >= nb) { const int real = a >= nb ? c : a; items.push_back((unsigned short)((real << 8) | real)); diag_done[real] = 1; }
      else items.push_back((unsigned short)((std::min(a, c) << 8) | std::max(a, c)));
    };
    add(nbe - 1, r);
    for (int k = 1; k < nbe / 2; ++k) add((r + k) % (nbe - 1), (r - k + nbe - 1) % (nbe - 1));
    rounds.push_back(items);
  }
  std::vector<unsigned short> cur;
  for (int i = 0; i < nb; ++i)
    if (!diag_done[i]) {
      cur.push_back((unsigned short)((i << 8) | i));
      if ((int)cur.size() == kFusedWarps) { rounds.push_back(cur); cur.clear(); }
    }
  if (!cur.empty()) rounds.push_back(cur);
  std::vector<unsigned short> flat;
  int nr = 0;
  for (const auto& rd : rounds)
    for (size_t o = 0; o < rd.size(); o += kFusedWarps, ++nr)
      for (int w = 0; w < kFusedWarps; ++w) flat.push_back(o + w < rd.size() ? rd[o + w] : kSchedIdle);
  *nrounds = nr;
  return flat;
}

}  // namespace


namespace {
// dz = sum_b dzk[b];  dgamma[b] / dalpha[b] = fixed-order sums of the row-block partials of zchain_kernel (+ the direct term
// -1/2 n_b beta N of dalpha)
__global__ void bound_fin_kernel(const double* dzk, const double* dad, const double* zpart, const double* alpha, double* dz,
                                 double* dgamma, double* dalpha, int b_count, int mq, int q, int nrb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < mq) {
    double s = 0;
    for (int b = 0; b < b_count; ++b) s += dzk[(size_t)b * mq + i];
    dz[i] = s;
  } else if (i < mq + b_count * (q + 1)) {
    const int j = i - mq, b = j / (q + 1), k = j - b * (q + 1);
    double s = 0;
    for (int r = 0; r < nrb; ++r) s += zpart[((size_t)b * nrb + r) * (q + 1) + k];
    if (k < q) dgamma[b * q + k] = s;
    else dalpha[b] = s / alpha[b] + dad[b];
  }
}
// final fixed-order sums of the chain partials
__global__ void chain_reduce_kernel(const double* dzp, const double* dgp, const double* dap, const double* dzd,
                                    double* dz, double* dgamma, double* dalpha, int grid, int b_count, int m, int mp,
                                    int q, int qp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nz = m * q, ng = b_count * q;
  if (i < nz) {
    const int mm_ = i / q, qq = i % q;
    double s = 0;
    for (int c = 0; c < grid; ++c)
      for (int b = 0; b < b_count; ++b) s += dzp[((size_t)c * b_count + b) * mp * qp + mm_ * qp + qq];
    for (int b = 0; b < b_count; ++b) s += dzd[(size_t)b * nz + i];
    dz[i] = s;
  } else if (i < nz + ng) {
    const int j = i - nz, b = j / q, qq = j % q;
    double s = 0;
    for (int c = 0; c < grid; ++c) s += dgp[((size_t)c * b_count + b) * qp + qq];
    dgamma[j] = s;
  } else if (i < nz + ng + b_count) {
    const int b = i - nz - ng;
    double s = 0;
    for (int c = 0; c < grid; ++c) s += dap[(size_t)c * b_count + b];
    dalpha[b] = s;
  }
}
}  // namespace

namespace {
void fine_mark(dpgp_handle* h, const char* name) {
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, h->fine_stream);
  h->fine_ev.emplace_back(name, e);
}
}  // namespace
namespace {
int launch_factor(dpgp_handle* h, const FactorParams& f, cudaStream_t st) {
  if (h->m <= kFacMaxM && !h->force_global_factor) factor_kernel<<<h->b, kFacThreads, fac_smem_bytes(h->m), st>>>(f);
  else { FactorGlobalParams g{f, h->fwork, h->ftmp}; factor_global_kernel<<<h->b, 512, 0, st>>>(g); }
  POST_LAUNCH(h, "factor_kernel");
  return DPGP_OK;
}
// C (n x m) = alpha * op(A) diag(w) op(B); strides in elements (see factor.cuh: MmJob)
MmJob mm_job(const double* a, long long sa, int ai, int ak, const double* bm, long long sb, int bk, int bj, double* c, long long sc,
             int ldc, int n, int m, int k, int flags, const double* w = nullptr, long long sw = 0, int wk = 1) {
  MmJob j{};
  j.a = a; j.bm = bm; j.c = c; j.w = w; j.sa = sa; j.sb = sb; j.sc = sc; j.sw = sw;
  j.ai = ai; j.ak = ak; j.bk = bk; j.bj = bj; j.ldc = ldc; j.wk = wk; j.n = n; j.m = m; j.k = k; j.flags = flags; j.alpha = 1.0;
  return j;
}
// X = L^-1 B for up to two right-hand sides per launch; B(i,j) = src[i*si + j*sj] (nc columns), X row-major with ldo
TrsmJob ts_job(const double* l, long long sl, const double* src, long long ss, int si, int sj, double* out, long long so, int ldo, int nc) {
  TrsmJob j{};
  j.l = l; j.src = src; j.out = out; j.sl = sl; j.ss = ss; j.so = so; j.si = si; j.sj = sj; j.ldo = ldo; j.nc = nc;
  return j;
}
int launch_trsm(dpgp_handle* h, std::initializer_list<TrsmJob> list, cudaStream_t st) {
  TrsmJobs js{};
  js.m = h->m;
  int blocks = 0;
  for (const TrsmJob& j : list) {
    TrsmJob& d = js.j[js.count++];
    d = j; d.blk0 = blocks;
    blocks += (j.nc + kTsCols - 1) / kTsCols;
  }
  trsm_cols_kernel<<<dim3(blocks, h->b), 128, ts_smem_bytes(h->m), st>>>(js);
  POST_LAUNCH(h, "trsm_cols_kernel");
  return DPGP_OK;
}
int launch_jobs(dpgp_handle* h, std::initializer_list<MmJob> list, cudaStream_t st) {
  MmJobs js{};
  int tiles = 0;
  for (const MmJob& j : list) {
    MmJob& d = js.j[js.count++];
    d = j;
    d.tiles_n = (j.n + kMmTile - 1) / kMmTile; d.tiles_m = (j.m + kMmTile - 1) / kMmTile; d.tile0 = tiles;
    tiles += d.tiles_n * d.tiles_m;
  }
  mm_jobs_kernel<<<dim3(tiles, h->b), 128, 0, st>>>(js);
  POST_LAUNCH(h, "mm_jobs_kernel");
  return DPGP_OK;
}
// L = chol(K_uu + 1e-8 I) and Linv = L^-1: everything of the chain that needs only (Z, gamma, alpha)
int kuu_factor(dpgp_handle* h, const double* z, const double* gamma, const double* alpha, cudaStream_t st) {
  FactorParams f{};
  f.z = z; f.gamma = gamma; f.alpha = alpha; f.out = h->linv; f.lout = h->lk; f.scal = nullptr; f.bad = h->bad; f.mode = 0; f.m = h->m; f.q = h->q;
  f.bad_offset = 0;
  return launch_factor(h, f, st);
}
void launch_zchain(dpgp_handle* h, const ZChainParams& zc, cudaStream_t st) {
  const size_t smem = (size_t)h->m * h->q * sizeof(double);      // <= 256 * 32 * 8 = 64 KB (opt-in in dpgp_create)
  const dim3 grid((h->m + kZcRows - 1) / kZcRows, h->b);
  if (h->q <= 16) zchain_kernel<16><<<grid, kZcRows * 32, smem, st>>>(zc);
  else zchain_kernel<32><<<grid, kZcRows * 32, smem, st>>>(zc);
}
}  // namespace

extern "C" {

int dpgp_create(dpgp_handle** out, int device, int64_t n_local, int d, int q, int m, int b, int mode,
                const dpgp_options* opt) {
  if (!out) return DPGP_E_ARG;
  *out = nullptr;
  dpgp_handle* h = new dpgp_handle();
  *out = h;      // returned even on failure so that dpgp_last_error() can be read; caller destroys it
  h->guard = getenv("DPGP_GUARD") != nullptr;
  if (n_local < 1 || d < 1 || q < 1 || q > kMaxQ || m < 1 || m > kMaxM || b < 1 || (mode != 0 && mode != 1))
    return fail(h, DPGP_E_ARG, "bad shape: n=%lld d=%d q=%d (1..%d) m=%d (1..%d) b=%d mode=%d", (long long)n_local, d, q,
                kMaxQ, m, kMaxM, b, mode);
  if (mode == DPGP_MODE_D && b != d) return fail(h, DPGP_E_ARG, "D-mode needs kernel batch b == d (%d != %d)", b, d);
  {
    int count = 0;
    CU(h, cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return fail(h, DPGP_E_ARG, "device %d does not exist (%d visible)", device, count);
  }
  DeviceGuard guard(device);
  cudaDeviceProp prop; CU(h, cudaGetDeviceProperties(&prop, device));
  h->device = device; h->sms = prop.multiProcessorCount;
  h->n = n_local; h->d = d; h->q = q; h->qp = pad_q(q); h->m = m; h->mp = round_up(m, 8); h->mt = (m + 1) / 2;
  h->t2 = h->mt * (h->mt + 1) / 2; h->b = b; h->mode = mode;
  h->ncols = (mode == DPGP_MODE_T) ? d : 1; h->cpad = h->ncols;
  h->expv = (opt && opt->exp_variant) ? opt->exp_variant : 4;
  if (h->expv < 1 || h->expv > 6) return fail(h, DPGP_E_ARG, "exp_variant must be 0..6");
  if (!kExperimental && h->expv != 1 && h->expv != 4)
    return fail(h, DPGP_E_ARG, "exp_variant %d is an experimental variant: rebuild with `make EXPERIMENTAL=1` (default build: 0/4 and 1)", h->expv);
  h->expv_fwd = (h->expv == 4 && !getenv("DPGP_FWD_TABLE256")) ? 8 : h->expv;
  h->bwd_variant = (opt && opt->bwd_variant) ? opt->bwd_variant : 6;
  if (h->bwd_variant < 1 || h->bwd_variant > 8) return fail(h, DPGP_E_ARG, "bwd_variant must be 0..8");
  if (!kExperimental && h->bwd_variant != 1 && h->bwd_variant < 6)
    return fail(h, DPGP_E_ARG, "bwd_variant %d is an experimental variant: rebuild with `make EXPERIMENTAL=1` (default build: 0, 1, 6, 7, 8)", h->bwd_variant);
  h->chain_variant = (opt && opt->chain_variant) ? opt->chain_variant : 1;
  if (h->chain_variant < 1 || h->chain_variant > 2) return fail(h, DPGP_E_ARG, "chain_variant must be 0..2");
  if (!kExperimental && h->chain_variant != 1)
    return fail(h, DPGP_E_ARG, "chain_variant 2 is an experimental variant: rebuild with `make EXPERIMENTAL=1`");
  h->grid = (opt && opt->max_ctas > 0) ? opt->max_ctas : h->sms;
  const size_t smem_cap = prop.sharedMemPerBlockOptin;
  // ---- psi2 forward configuration: consumer threads TC (multiple of 32) minimising idle tile slots; +1 producer warp.
  // The per-thread accumulators of every pass live in shared memory; for M > ~200 they do not fit next to a useful row tile,
  // so the pair triangle is split over f_splits launches of f_t2 tiles each.
  h->f_chunk = (opt && opt->psi2_chunk > 0) ? opt->psi2_chunk : 32;
  int tc = 0;
  for (h->f_splits = 1; h->f_splits <= 8; ++h->f_splits) {
    h->f_t2 = (h->t2 + h->f_splits - 1) / h->f_splits;
    // registers are allocated per 4 warps: 12 warps (11 consumers + producer) leave 168 registers per thread
    if (opt && opt->psi2_threads > 0) tc = std::min(352, round_up(opt->psi2_threads, 32));
    else {
      double best = 1e30;
      for (int t = 352; t >= 96; t -= 32) {
        int np = (h->f_t2 + t - 1) / t;
        double waste = (double)(np * t - h->f_t2) / (np * t) + (t < 256 ? 0.05 : 0.0);
        if (waste < best - 1e-9) { best = waste; tc = t; }
      }
    }
    h->f_threads = tc + 32;
    h->f_npass = (h->f_t2 + tc - 1) / tc;
    auto fsm = [&](int chunk) {
      return ((size_t)h->f_npass * tc * 4 + kStages * (size_t)chunk * (h->mp + h->qp) + 2 * (size_t)h->mt * h->qp) * 8 +
             (((size_t)h->f_npass * tc + 1) & ~(size_t)1) * 4 + 2 * kStages * 8 + (h->expv_fwd == 8 ? kExpTabSizeFwd : kExpTabSize) * 8;
    };
    int chunk = h->f_chunk;
    const int min_chunk = h->f_splits < 8 ? 16 : 4;       // rather split the triangle than starve the row tile
    while (chunk > min_chunk && fsm(chunk) > smem_cap) chunk /= 2;
    h->f_smem = fsm(chunk);
    if (h->f_smem <= smem_cap) { h->f_chunk = chunk; break; }
  }
  if (h->f_smem > smem_cap) return fail(h, DPGP_E_ARG, "psi2 forward needs %zu B of shared memory (> %zu)", h->f_smem, smem_cap);

  // ---- psi2 backward, pair side: consumer threads own one half tile each
  {
    double best = 1e30; int ptc = 0;
    for (int t = 352; t >= 96; t -= 32) {
      int jb = (2 * h->t2 + t - 1) / t;
      double waste = (double)(jb * t - 2 * h->t2) / (jb * t) + (double)(h->grid % jb) / h->grid;
      if (jb <= h->grid && waste < best - 1e-9) { best = waste; ptc = t; }
    }
    if (!ptc) ptc = 352;
    h->p_threads = ptc + 32;
    h->p_jb = (2 * h->t2 + ptc - 1) / ptc;
    h->p_ng = std::max(1, h->grid / h->p_jb);
    auto psm = [&](int chunk) { return (kStages * (size_t)chunk * (h->mp + h->qp) + 2 * (size_t)h->mt * h->qp) * 8 + 2 * kStages * 8 + kExpTabSize * 8; };
    while (h->p_chunk > 4 && psm(h->p_chunk) > smem_cap) h->p_chunk /= 2;
    h->p_smem = psm(h->p_chunk);
  }
  // ---- psi2 backward, n side
  {
    int t = (int)((smem_cap - 1024) / ((size_t)h->mp * 8)) / 32 * 32;
    h->n_threads = std::max(32, std::min(192, t));
    h->n_smem = (size_t)h->mp * h->n_threads * 8 + kExpTabSize * 8;
  }
  // ---- psi2 backward, fused
  h->k = launchers_for(h->qp);
  if (!h->k) return fail(h, DPGP_E_ARG, "no kernels were built for the padded latent dimension %d", h->qp);
  std::vector<unsigned short> sched = build_fused_schedule(h->mp / 8, &h->u_nrounds);
  {
    h->u_rows = 2;
    h->u_smem = h->k->fused_smem(2, h->mp);
    if (const char* e = getenv("DPGP_U_ROWS")) { if (atoi(e) == 1) h->u_smem = smem_cap + 1; }   // development switch
    if (h->u_smem > smem_cap || h->qp > 12) { h->u_rows = 1; h->u_smem = h->k->fused_smem(1, h->mp); }
#ifdef DPGP_EXPERIMENTAL
    if (h->bwd_variant == 5) {                           // warp-specialised: 64-row groups, QP <= 12, must fit
      if (h->u_rows == 2 && h->qp <= 12 && h->k->ws_smem(h->mp) <= smem_cap) h->u_smem = h->k->ws_smem(h->mp);
      else h->bwd_variant = 1;
    }
    if (h->bwd_variant == 4) {
      if (h->k->fused2_smem(h->mp) <= smem_cap) { h->u_rows = 1; h->u_smem = h->k->fused2_smem(h->mp); }
      else h->bwd_variant = 1;                            // does not fit (M > 128 or so): single-team kernel
    }
#endif
    if (h->bwd_variant == 7) {                           // tensor-core slice products: 64-row groups, QP <= 16, Mp <= 128 (shared memory)
      h->um_smem = umma_smem_bytes(h->mp, h->qp);
      if (h->qp > 16 || h->um_smem > smem_cap)
        return fail(h, DPGP_E_ARG, "bwd_variant 7 needs Q <= 16 and M <= 128 (%zu B of shared memory > %zu)", h->um_smem, smem_cap);
      h->u_rows = 2; h->u_smem = 0;
    }
    if (h->bwd_variant == 8) {                           // DMMA contractions: 64-row groups, 8 <= QP <= 12, must fit
      const size_t need = h->k->mma_smem(h->mp);
      if (h->u_rows != 2 || need == 0 || need > smem_cap)
        return fail(h, DPGP_E_ARG, "bwd_variant 8 needs 8 <= padded Q <= 12 and %zu B of shared memory (<= %zu)", need, smem_cap);
      h->u_smem = need;
    }
    if (h->u_smem > smem_cap) return fail(h, DPGP_E_ARG, "fused psi2 backward needs %zu B of shared memory (> %zu)", h->u_smem, smem_cap);
    const int64_t ngroups = cdiv64(n_local, 32 * h->u_rows);
    h->u_grid = (int)std::min<int64_t>(ngroups * b, (int64_t)h->grid);
    const int64_t per = cdiv64(ngroups * b, h->u_grid);
    h->u_nseg = (int)std::min<int64_t>(b, cdiv64(per, ngroups) + 1);
    h->u_slice = (size_t)h->u_nrounds * kFusedWarps * 64 * h->qp;
    if (h->bwd_variant >= 6) { h->u_nseg = 1; h->u_slice = (size_t)kFusedWarps * 2 * h->mp * h->qp; }   // per-warp dz slices
  }
  const int pgrid = h->p_jb * h->p_ng;
  // a CTA works on a contiguous range of (cluster, chunk) items: number of distinct clusters it can meet
  auto nseg_for = [&](int64_t nchunks, int workers) {
    const int64_t per = cdiv64(nchunks * b, workers);
    return (int)std::min<int64_t>(b, cdiv64(per, nchunks) + 1);
  };
  h->f_nseg = nseg_for(cdiv64(n_local, h->f_chunk), h->grid);
  h->p1_grid = 2 * h->grid;
  h->p1_nseg = nseg_for(cdiv64(n_local, kP1Rows), h->p1_grid);
  h->p_nseg = nseg_for(cdiv64(n_local, h->p_chunk), h->p_ng);
  h->cs_grid = (int)std::min<int64_t>(h->grid, std::max<int64_t>(1, n_local / 16));

  // ---- psi1 backward + chain
#ifdef DPGP_EXPERIMENTAL
  if (h->bwd_variant == 3 && (h->u_rows != 2 || h->qp > 12)) h->bwd_variant = 1;      // tensor-core variant: 64-row groups, QP <= 12
  if (h->bwd_variant == 5) {
    Psi2BwdFusedParams dummy{};
    if (!h->k->psi2_bwd_ws(h->expv, 0, h->u_smem, nullptr, dummy, true)) return fail(h, DPGP_E_CUDA, "cannot configure the warp-specialised fused backward");
  }
  if (h->bwd_variant == 4) {
    Psi2BwdFusedParams dummy{};
    if (!h->k->psi2_bwd_fused2(h->expv, 0, h->u_smem, nullptr, dummy, true)) return fail(h, DPGP_E_CUDA, "cannot configure the two-team fused backward");
  }
  if (h->bwd_variant == 3) {
    Psi2BwdFusedParams dummy{};
    if (!h->k->psi2_bwd_tc(h->expv, 0, h->u_smem, nullptr, dummy, true)) return fail(h, DPGP_E_CUDA, "cannot configure psi2_bwd_tc_kernel");
  }
#endif
  // 16-row tiles let two CTAs share an SM (measured 12.8 vs 14.2 ms at 262 144 rows: the kernel is latency-bound across
  // its phases, so a second CTA fills the gaps); 32-row tiles otherwise.  DPGP_C2_ROWS = 16 / 32 overrides (development).
  h->c2_rows = 16; h->c2_smem = h->k->chain2_smem(16, h->mp);
  {
    const char* e = getenv("DPGP_C2_ROWS");
    const bool two_fit = 2 * h->c2_smem + 4096 <= smem_cap;
    if ((e && atoi(e) == 32) || (!e && !two_fit)) {
      const size_t s32 = h->k->chain2_smem(32, h->mp);
      if (s32 <= smem_cap) { h->c2_rows = 32; h->c2_smem = s32; }
    }
  }
  h->c2_grid = (int)std::min<int64_t>(cdiv64(n_local, h->c2_rows), (int64_t)h->grid * (h->c2_smem * 2 + 4096 <= smem_cap ? 2 : 1));
  // few row tiles (small N): split the clusters over groups of CTAs as well, up to ~2 CTAs per SM in total
  h->c2_groups = (int)std::max<int64_t>(1, std::min<int64_t>(b, (2 * (int64_t)h->grid) / std::max(1, h->c2_grid)));
  if ((int64_t)n_local * q * h->c2_groups > ((int64_t)1 << 24)) h->c2_groups = 1;      // partial buffers only where they are small
  const int cgmax = std::max(h->grid, h->c2_grid * h->c2_groups);
  // ---- workspace
  const size_t bn = (size_t)b * (size_t)n_local, mm = (size_t)m * m, mc = (size_t)m * h->ncols;
  int rc;
  if ((rc = ws_alloc(h, &h->r, bn * h->mp))) return rc;
  if ((rc = ws_alloc(h, &h->v, bn * h->qp))) return rc;
  if ((rc = ws_alloc(h, &h->bco, h->chain_variant == 2 ? bn * h->mp : 1))) return rc;
  if ((rc = ws_alloc(h, &h->dv, bn * h->qp))) return rc;
  if ((rc = ws_alloc(h, &h->f_part, (size_t)h->grid * h->f_nseg * h->f_npass * (h->f_threads - 32) * 4))) return rc;
  if ((rc = ws_alloc(h, &h->f_tags, (size_t)h->grid * h->f_nseg))) return rc;
  if ((rc = ws_alloc(h, &h->p1_part, (size_t)h->p1_grid * h->p1_nseg * h->mp * h->cpad))) return rc;
  if ((rc = ws_alloc(h, &h->p1_tags, (size_t)h->p1_grid * h->p1_nseg))) return rc;
  if ((rc = ws_alloc(h, &h->cs_part, (size_t)h->cs_grid * (d + 2)))) return rc;
  if ((rc = ws_alloc(h, &h->bp_part, h->bwd_variant == 2 ? (size_t)pgrid * h->p_nseg * (h->p_threads - 32) * 2 * h->qp : 1))) return rc;
  if ((rc = ws_alloc(h, &h->bp_tags, (size_t)pgrid * h->p_nseg))) return rc;
  if ((rc = ws_alloc(h, &h->ddsym, (size_t)b * mm * h->qp))) return rc;
  if ((rc = ws_alloc(h, &h->bad, (size_t)b))) return rc;
  {
    double** mats[] = {&h->lk, &h->la, &h->linv, &h->kinv, &h->t1, &h->hmat, &h->lainv, &h->rmat, &h->smat, &h->gmat, &h->rtg, &h->gtg, &h->wmat};
    for (double** pm : mats) if ((rc = ws_alloc(h, pm, (size_t)b * mm))) return rc;
    double** cols[] = {&h->cm, &h->umat, &h->pu, &h->c1};
    for (double** pc : cols) if ((rc = ws_alloc(h, pc, (size_t)b * mc))) return rc;
    if ((rc = ws_alloc(h, &h->fscal, (size_t)b * 4))) return rc;
    h->force_global_factor = getenv("DPGP_FORCE_GLOBAL_FACTOR") != nullptr;      // development: exercise the M > 144 path at small M
    if (m > kFacMaxM || h->force_global_factor) {
      if ((rc = ws_alloc(h, &h->fwork, (size_t)b * mm))) return rc;
      if ((rc = ws_alloc(h, &h->ftmp, (size_t)b * mm))) return rc;
    }
    CU(h, cudaFuncSetAttribute(factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fac_smem_bytes(kFacMaxM)));
    CU(h, cudaFuncSetAttribute(trsm_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts_smem_bytes(kMaxM)));
    if (!getenv("DPGP_NO_SIDE_STREAM")) {
      CU(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
      CU(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
      CU(h, cudaEventCreateWithFlags(&h->ev_kuu, cudaEventDisableTiming));
    }
  }
  if ((rc = ws_alloc(h, &h->fb, (size_t)b))) return rc;
  if ((rc = ws_alloc(h, &h->zpart, (size_t)b * ((m + kZcRows - 1) / kZcRows) * (q + 1)))) return rc;
  if ((rc = ws_alloc(h, &h->dk, (size_t)b * mm))) return rc;
  if ((rc = ws_alloc(h, &h->dzk, (size_t)b * m * q))) return rc;
  if ((rc = ws_alloc(h, &h->dzd, (size_t)b * m * q))) return rc;
  if ((rc = ws_alloc(h, &h->dadirect, (size_t)b))) return rc;
  if ((rc = ws_alloc(h, &h->dzp, (size_t)cgmax * (h->chain_variant == 2 ? b : 1) * h->mp * h->qp))) return rc;
  if ((rc = ws_alloc(h, &h->dgp, (size_t)cgmax * b * h->qp))) return rc;
  if ((rc = ws_alloc(h, &h->dap, (size_t)cgmax * b))) return rc;
  if (h->c2_groups > 1) {
    if ((rc = ws_alloc(h, &h->c2_mu_part, (size_t)h->c2_groups * n_local * q))) return rc;
    if ((rc = ws_alloc(h, &h->c2_s_part, (size_t)h->c2_groups * n_local * q))) return rc;
  }
  if ((rc = ws_alloc(h, &h->dummy, (size_t)b * (q + 1) + 16))) return rc;
#ifdef DPGP_EXPERIMENTAL
  {
    const size_t nblk = nside_num_blocks(h->mp);
    if ((rc = ws_alloc(h, &h->dtab, nblk * 32 * h->qp))) return rc;
    if ((rc = ws_alloc(h, &h->gtab, (size_t)b * nblk * 32))) return rc;
  }
#endif
  if ((rc = ws_alloc(h, &h->exptab, (size_t)kExpTabSize))) return rc;
  if ((rc = ws_alloc(h, &h->exptab_fwd, (size_t)kExpTabSizeFwd))) return rc;
  if ((rc = ws_alloc(h, &h->u_sched, sched.size()))) return rc;
  if (h->bwd_variant != 2) {
    if ((rc = ws_alloc(h, &h->u_part, (size_t)h->u_grid * h->u_nseg * h->u_slice))) return rc;
    if ((rc = ws_alloc(h, &h->u_tags, (size_t)h->u_grid * h->u_nseg))) return rc;
  }
  if (h->bwd_variant == 7) {
    if ((rc = ws_alloc(h, &h->um_wtab, (size_t)b * h->u_nrounds * 512))) return rc;
    if ((rc = ws_alloc(h, &h->um_dprime, (size_t)b * h->u_nrounds * 8 * kUmDStage))) return rc;
    if ((rc = ws_alloc(h, &h->um_scale, (size_t)b))) return rc;
  }
  {
    double tab[kExpTabSize];
    const int tsize = 1 << exp_tab_bits(h->expv);      // entries actually indexed by this variant; the rest repeat
    for (int j = 0; j < kExpTabSize; ++j) tab[j] = (double)exp2l((long double)(j % tsize) / (long double)tsize);
    CU(h, cudaMemcpy(h->exptab, tab, sizeof tab, cudaMemcpyHostToDevice));
    std::vector<double> tabf(kExpTabSizeFwd);
    for (int j = 0; j < kExpTabSizeFwd; ++j) tabf[j] = (double)exp2l((long double)j / (long double)kExpTabSizeFwd);
    CU(h, cudaMemcpy(h->exptab_fwd, tabf.data(), sizeof(double) * kExpTabSizeFwd, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->u_sched, sched.data(), sched.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
  }
  CU(h, cudaMemset(h->bad, 0, sizeof(int) * b));
  for (int i = 0; i < kNumPhases; ++i) { CU(h, cudaEventCreate(&h->ev0[i])); CU(h, cudaEventCreate(&h->ev1[i])); }

  // ---- opt in to large dynamic shared memory for every instantiation that can be selected
  const size_t p1_smem = ((size_t)kP1Rows * h->mp + (size_t)kP1Cols * kP1Rows) * 8;
  CU(h, h->k->cfg_smem(h->expv, h->f_smem, p1_smem, h->u_rows, h->u_smem));
  if (h->bwd_variant == 8 && !h->k->psi2_bwd_mma(h->expv, 0, h->u_smem, nullptr, Psi2BwdFusedParams{}, true))
    return fail(h, DPGP_E_ARG, "bwd_variant 8 is not available for the padded latent dimension %d / exp_variant %d (built for the default exp only)", h->qp, h->expv);
  if (h->bwd_variant == 7 && !h->k->psi2_bwd_umma(h->expv, 0, h->um_smem, nullptr, Psi2BwdUmmaParams{}, true))
    return fail(h, DPGP_E_ARG, "bwd_variant 7 is not available for the padded latent dimension %d / exp_variant %d (built for the default exp only)", h->qp, h->expv);
#ifdef DPGP_EXPERIMENTAL
  {
    const size_t g1_smem = p1_smem + (size_t)kP1Cols * h->mp * 8;
    const size_t ch_smem = (2 * (size_t)kChRows * h->mp + (size_t)h->mp * h->qp) * 8;
    CU(h, h->k->cfg_smem_x(h->expv, h->p_smem, h->n_smem, g1_smem, ch_smem));
  }
#endif
  CU(h, h->k->chain2_cfg(h->c2_rows, h->c2_smem));
  CU(h, cudaFuncSetAttribute(zchain_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxM * 16 * 8));
  CU(h, cudaFuncSetAttribute(zchain_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxM * 32 * 8));
  CU(h, cudaFuncSetAttribute(small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes(kSmallMaxT, kMaxQ)));
  CU(h, cudaFuncSetAttribute(small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem_bytes(kSmallMaxT, kMaxQ)));
  return DPGP_OK;
}

int dpgp_destroy(dpgp_handle* h) {
  if (!h) return DPGP_OK;
  DeviceGuard guard(h->device);
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_kuu) cudaEventDestroy(h->ev_kuu);
  for (void* p : h->allocs) cudaFree(p);
  for (int i = 0; i < kNumPhases; ++i) { if (h->ev0[i]) cudaEventDestroy(h->ev0[i]); if (h->ev1[i]) cudaEventDestroy(h->ev1[i]); }
  delete h;
  return DPGP_OK;
}

int dpgp_check_guards(dpgp_handle* h) {
  if (!h) return DPGP_E_ARG;
  if (!h->guard) return fail(h, DPGP_E_ARG, "the handle was created without DPGP_GUARD in the environment");
  DeviceGuard dg(h->device);
  CU(h, cudaDeviceSynchronize());
  std::vector<unsigned char> host;
  int bad = 0;
  for (size_t i = 0; i < h->guarded.size(); ++i) {
    const auto& g = h->guarded[i];
    const size_t upper = g.rounded - g.bytes + kGuardBytes;
    host.resize(kGuardBytes + upper);
    CU(h, cudaMemcpy(host.data(), g.base, kGuardBytes, cudaMemcpyDeviceToHost));
    CU(h, cudaMemcpy(host.data() + kGuardBytes, g.base + kGuardBytes + g.bytes, upper, cudaMemcpyDeviceToHost));
    for (size_t k = 0; k < host.size(); ++k)
      if (host[k] != kGuardPattern) {
        if (!bad) fail(h, DPGP_E_CUDA, "guard band of workspace buffer %zu (%zu bytes) overwritten %s it, %zu bytes from the buffer",
                       i, g.bytes, k < kGuardBytes ? "below" : "above", k < kGuardBytes ? kGuardBytes - k : k - kGuardBytes + 1);
        ++bad;
        break;
      }
  }
  return bad;
}

int dpgp_has_experimental(void) { return kExperimental ? 1 : 0; }
int dpgp_limits(int* max_q, int* max_m) { if (max_q) *max_q = kMaxQ; if (max_m) *max_m = kMaxM; return DPGP_OK; }

const char* dpgp_last_error(const dpgp_handle* h) { return h ? h->err.c_str() : "null handle"; }
size_t dpgp_stats_len(const dpgp_handle* h) {
  return (size_t)h->b * h->m * h->m + (size_t)h->b * h->m * h->ncols + h->d + 2;
}
size_t dpgp_workspace_bytes(const dpgp_handle* h) { return h->ws_bytes; }
int64_t dpgp_launch_count(const dpgp_handle* h) { return h->launches; }

int dpgp_check(dpgp_handle* h, void* stream) {
  if (!h) return DPGP_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CU(h, cudaStreamSynchronize(st));
  std::vector<int> bad(h->b);
  CU(h, cudaMemcpy(bad.data(), h->bad, sizeof(int) * h->b, cudaMemcpyDeviceToHost));
  for (int b = 0; b < h->b; ++b)
    if (bad[b]) {
      CU(h, cudaMemset(h->bad, 0, sizeof(int) * h->b));
      const bool second = bad[b] > 1000;
      return fail(h, DPGP_E_NOT_PD, "Cholesky of %s met a non-positive pivot at row %d for kernel-batch entry %d",
                  second ? "beta*H + I" : "K_uu + 1e-8 I", (bad[b] % 1000) - 1, b);
    }
  return DPGP_OK;
}

namespace {
// One Adam update of a flat parameter tensor, TensorFlow-1 formulation (tf.train.AdamOptimizer, the optimiser of
// every reference script, e.g. test/synthetic_data_hard_test.py:143):
//   lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t);  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//   theta -= lr_t m / (sqrt(v) + eps)
// The step count t lives on the device so that a captured CUDA graph of a training iteration can be replayed.
// HBM-bound: 56 bytes per element (read theta, g, m, v; write theta, m, v), double2 accesses, grid-stride.
__global__ void __launch_bounds__(256) adam_kernel(double* __restrict__ theta, const double* __restrict__ g, double* __restrict__ m,
                                                   double* __restrict__ v, int64_t n, const int64_t* __restrict__ step,
                                                   double lr, double b1, double b2, double eps) {
  const double t = (double)*step;
  const double lr_t = lr * sqrt(1.0 - pow(b2, t)) / (1.0 - pow(b1, t));
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(theta) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  auto upd = [&](double& th, double gg, double& mm_, double& vv) {
    mm_ = fma(b1, mm_, (1.0 - b1) * gg);
    vv = fma(b2, vv, (1.0 - b2) * gg * gg);
    th -= lr_t * mm_ / (sqrt(vv) + eps);
  };
  if (vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
      double2 th = reinterpret_cast<double2*>(theta)[i], mm_ = reinterpret_cast<double2*>(m)[i], vv = reinterpret_cast<double2*>(v)[i];
      const double2 gg = __ldcs(reinterpret_cast<const double2*>(g) + i);
      upd(th.x, gg.x, mm_.x, vv.x); upd(th.y, gg.y, mm_.y, vv.y);
      reinterpret_cast<double2*>(theta)[i] = th; reinterpret_cast<double2*>(m)[i] = mm_; reinterpret_cast<double2*>(v)[i] = vv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) upd(theta[n - 1], g[n - 1], m[n - 1], v[n - 1]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) upd(theta[i], g[i], m[i], v[i]);
  }
}
}  // namespace

int dpgp_adam(dpgp_handle* h, double* d_param, const double* d_grad, double* d_m, double* d_v, int64_t n,
              const int64_t* d_step, double lr, double beta1, double beta2, double eps, void* stream) {
  if (!h || !d_param || !d_grad || !d_m || !d_v || !d_step || n < 0) return fail(h, DPGP_E_ARG, "dpgp_adam: null argument");
  if (n == 0) return DPGP_OK;
  const int grid = (int)std::min<int64_t>((n / 2 + 255) / 256 + 1, (int64_t)h->sms * 8);
  adam_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_param, d_grad, d_m, d_v, n, d_step, lr, beta1, beta2, eps);
  POST_LAUNCH(h, "adam_kernel");
  return DPGP_OK;
}

namespace {
constexpr int kAdamMaxTensors = 16;
struct AdamMultiParams {
  double* theta[kAdamMaxTensors]; const double* g[kAdamMaxTensors]; double* m[kAdamMaxTensors]; double* v[kAdamMaxTensors];
  int64_t n[kAdamMaxTensors];
  const int64_t* step; double lr, b1, b2, eps;
  double gs[kAdamMaxTensors];                // gradient used = gs * g (* sigmoid(raw) where raw != nullptr); dpgp_adam_multi: 1, nullptr
  const double* raw[kAdamMaxTensors];
};
// All trainable tensors of a model in one launch (blockIdx.y = tensor): at the reference's problem sizes the eleven
// separate launches were 10 % of a training iteration.  Same arithmetic as adam_kernel, element by element.
__global__ void __launch_bounds__(256) adam_multi_kernel(AdamMultiParams p) {
  const int k = blockIdx.y;
  const int64_t n = p.n[k];
  if ((int64_t)blockIdx.x * blockDim.x >= n) return;
  double* theta = p.theta[k]; const double* g = p.g[k]; double* m = p.m[k]; double* v = p.v[k];
  const double t = (double)*p.step;
  const double lr_t = p.lr * sqrt(1.0 - pow(p.b2, t)) / (1.0 - pow(p.b1, t));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double gg = __ldcs(g + i);
    if (p.raw[k]) gg *= 1.0 / (1.0 + exp(-p.raw[k][i]));          // chain of s = softplus(raw) (src/utils/types.py:52-57)
    gg *= p.gs[k];
    const double mm_ = fma(p.b1, m[i], (1.0 - p.b1) * gg);
    const double vv = fma(p.b2, v[i], (1.0 - p.b2) * gg * gg);
    m[i] = mm_; v[i] = vv;
    theta[i] -= lr_t * mm_ / (sqrt(vv) + p.eps);
  }
}
}  // namespace

namespace {
// objective = dp objective - hyper-prior - (f_hat - KL)   (dp_gp_lvm.py:148-154 / :670-676), and the step counter of the iteration
__global__ void train_pre_kernel(const double* scal, const double* gp, double* objective, int64_t* step) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { *objective = (scal[0] - scal[1]) - gp[0]; *step += 1; }
}
int adam_multi_impl(dpgp_handle* h, int count, double* const* d_params, const double* const* d_grads, double* const* d_ms,
                    double* const* d_vs, const int64_t* ns, const double* g_scale, const double* const* d_raws, const int64_t* d_step,
                    double lr, double beta1, double beta2, double eps, void* stream);
}  // namespace

int dpgp_adam_multi(dpgp_handle* h, int count, double* const* d_params, const double* const* d_grads, double* const* d_ms,
                    double* const* d_vs, const int64_t* ns, const int64_t* d_step, double lr, double beta1, double beta2, double eps,
                    void* stream) {
  return adam_multi_impl(h, count, d_params, d_grads, d_ms, d_vs, ns, nullptr, nullptr, d_step, lr, beta1, beta2, eps, stream);
}

int dpgp_train_tail(dpgp_handle* h, const double* d_scal, const double* d_gp, double* d_objective, int count, double* const* d_params,
                    const double* const* d_grads, double* const* d_ms, double* const* d_vs, const int64_t* ns, const double* g_scale,
                    const double* const* d_raws, int64_t* d_step, double lr, double beta1, double beta2, double eps, void* stream) {
  if (!h || !d_scal || !d_gp || !d_objective || !d_step || !g_scale) return fail(h, DPGP_E_ARG, "dpgp_train_tail: null argument");
  DeviceGuard guard(h->device);
  train_pre_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_scal, d_gp, d_objective, d_step);
  POST_LAUNCH(h, "train_pre_kernel");
  return adam_multi_impl(h, count, d_params, d_grads, d_ms, d_vs, ns, g_scale, d_raws, d_step, lr, beta1, beta2, eps, stream);
}

namespace {
int adam_multi_impl(dpgp_handle* h, int count, double* const* d_params, const double* const* d_grads, double* const* d_ms,
                    double* const* d_vs, const int64_t* ns, const double* g_scale, const double* const* d_raws, const int64_t* d_step,
                    double lr, double beta1, double beta2, double eps, void* stream) {
  if (!h || count < 0 || (count > 0 && (!d_params || !d_grads || !d_ms || !d_vs || !ns)) || !d_step)
    return fail(h, DPGP_E_ARG, "dpgp_adam_multi: null argument");
  for (int base = 0; base < count; base += kAdamMaxTensors) {
    AdamMultiParams p{};
    const int c = std::min(kAdamMaxTensors, count - base);
    int64_t nmax = 0;
    for (int i = 0; i < c; ++i) {
      if (ns[base + i] < 0 || (ns[base + i] > 0 && (!d_params[base + i] || !d_grads[base + i] || !d_ms[base + i] || !d_vs[base + i])))
        return fail(h, DPGP_E_ARG, "dpgp_adam_multi: null tensor %d", base + i);
      p.theta[i] = d_params[base + i]; p.g[i] = d_grads[base + i]; p.m[i] = d_ms[base + i]; p.v[i] = d_vs[base + i]; p.n[i] = ns[base + i];
      p.gs[i] = g_scale ? g_scale[base + i] : 1.0; p.raw[i] = d_raws ? d_raws[base + i] : nullptr;
      nmax = std::max(nmax, ns[base + i]);
    }
    if (nmax == 0) continue;
    p.step = d_step; p.lr = lr; p.b1 = beta1; p.b2 = beta2; p.eps = eps;
    const int gx = (int)std::min<int64_t>((nmax + 255) / 256, (int64_t)h->sms * 8);
    adam_multi_kernel<<<dim3(gx, c), 256, 0, (cudaStream_t)stream>>>(p);
    POST_LAUNCH(h, "adam_multi_kernel");
  }
  return DPGP_OK;
}
}  // namespace

namespace {
int small_fill(dpgp_handle* h, const dpgp_small_args* a, SmallParams& p, bool bwd) {
  if (!h || !a) return DPGP_E_ARG;
  const int T = a->truncation_level, mask = a->mask_size;
  if (T < 1 || T > kSmallMaxT || mask < 1 || h->d % mask != 0) return fail(h, DPGP_E_ARG, "dpgp_small: truncation_level must be 1..%d and mask_size must divide D", kSmallMaxT);
  if (h->mode == DPGP_MODE_T && h->b != T) return fail(h, DPGP_E_ARG, "dpgp_small: T-mode handle has B = %d, truncation_level = %d", h->b, T);
  if (!a->logits || !a->w1_raw || !a->w2_raw || !a->gamma_atoms_raw || !a->alpha_atoms_raw || !a->beta_atoms_raw || !a->phi ||
      (T > 1 && (!a->gamma1_raw || !a->gamma2_raw)))
    return fail(h, DPGP_E_ARG, "dpgp_small: null argument");
  if (!bwd && (!a->gamma || !a->alpha || !a->beta || !a->scal)) return fail(h, DPGP_E_ARG, "dpgp_small_fwd: null output");
  if (bwd && (!a->dgamma || !a->dalpha || !a->dbeta || (h->mode == DPGP_MODE_T && !a->dphi) || !a->dlogits || !a->dw1_raw || !a->dw2_raw ||
              !a->dgamma_atoms_raw || !a->dalpha_atoms_raw || !a->dbeta_atoms_raw || (T > 1 && (!a->dgamma1_raw || !a->dgamma2_raw))))
    return fail(h, DPGP_E_ARG, "dpgp_small_bwd: null argument");
  p.logits = a->logits; p.g1_raw = a->gamma1_raw; p.g2_raw = a->gamma2_raw; p.w1_raw = a->w1_raw; p.w2_raw = a->w2_raw;
  p.ga_raw = a->gamma_atoms_raw; p.aa_raw = a->alpha_atoms_raw; p.ba_raw = a->beta_atoms_raw;
  p.phi = a->phi; p.gamma = a->gamma; p.alpha = a->alpha; p.beta = a->beta; p.scal = a->scal;
  p.dphi = a->dphi; p.dgamma = a->dgamma; p.dalpha = a->dalpha; p.dbeta = a->dbeta; p.grad_out = a->grad_out;
  p.dlogits = a->dlogits; p.dg1_raw = a->dgamma1_raw; p.dg2_raw = a->dgamma2_raw; p.dw1_raw = a->dw1_raw; p.dw2_raw = a->dw2_raw;
  p.dga_raw = a->dgamma_atoms_raw; p.daa_raw = a->dalpha_atoms_raw; p.dba_raw = a->dbeta_atoms_raw;
  p.d = h->d; p.t = T; p.q = h->q; p.mask = mask; p.mode = h->mode == DPGP_MODE_T ? 0 : 1;
  p.s1 = a->alpha_prior_shape; p.s2 = a->alpha_prior_rate;
  return DPGP_OK;
}
}  // namespace

int dpgp_small_fwd(dpgp_handle* h, const dpgp_small_args* a, void* stream) {
  SmallParams p{};
  if (int rc = small_fill(h, a, p, false)) return rc;
  small_fwd_kernel<<<1, kSmallThreads, small_smem_bytes(p.t, p.q), (cudaStream_t)stream>>>(p);
  POST_LAUNCH(h, "small_fwd_kernel");
  return DPGP_OK;
}
int dpgp_small_bwd(dpgp_handle* h, const dpgp_small_args* a, void* stream) {
  SmallParams p{};
  if (int rc = small_fill(h, a, p, true)) return rc;
  small_bwd_kernel<<<1, kSmallThreads, small_smem_bytes(p.t, p.q), (cudaStream_t)stream>>>(p);
  POST_LAUNCH(h, "small_bwd_kernel");
  return DPGP_OK;
}

namespace {
__global__ void polygamma_kernel(const double* __restrict__ x, double* __restrict__ psi, double* __restrict__ tri, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    if (psi) psi[i] = v > 0.0 ? digamma_pos(v) : nan("");
    if (tri) tri[i] = v > 0.0 ? trigamma_pos(v) : nan("");
  }
}
}  // namespace

int dpgp_polygamma(const double* d_x, double* d_digamma, double* d_trigamma, int64_t n, void* stream) {
  if (n < 0) return DPGP_E_ARG;
  if (n == 0 || (!d_digamma && !d_trigamma)) return DPGP_OK;      // empty tensors (T = 1: no stick-breaking variables) are fine
  if (!d_x) return DPGP_E_ARG;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, 1184);
  polygamma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, d_digamma, d_trigamma, n);
  return cudaGetLastError() == cudaSuccess ? DPGP_OK : DPGP_E_CUDA;
}

int dpgp_fused_schedule(int num_mblocks, unsigned short* out, int cap) {
  if (num_mblocks < 1 || num_mblocks > kMaxM / 8) return DPGP_E_ARG;
  int nr = 0;
  const std::vector<unsigned short> s = build_fused_schedule(num_mblocks, &nr);
  if (out) for (size_t i = 0; i < s.size() && (int)i < cap; ++i) out[i] = s[i];
  return nr;
}

int dpgp_debug_launch_times(dpgp_handle* h, int enable, const char** names, float* us, int cap) {
  if (!h) return DPGP_E_ARG;
  int k = 0;
  if (!h->fine_ev.empty()) {
    cudaDeviceSynchronize();
    for (size_t i = 1; i < h->fine_ev.size(); ++i) {
      float ms = 0;
      if (k < cap && names && us && cudaEventElapsedTime(&ms, h->fine_ev[i - 1].second, h->fine_ev[i].second) == cudaSuccess) {
        names[k] = h->fine_ev[i].first; us[k] = ms * 1e3f; ++k;
      }
    }
    for (auto& pe : h->fine_ev) cudaEventDestroy(pe.second);
    h->fine_ev.clear();
  }
  h->fine = enable != 0;
  return k;
}

int dpgp_set_timing(dpgp_handle* h, int enabled) { if (!h) return DPGP_E_ARG; h->timing = enabled != 0; return DPGP_OK; }
int dpgp_get_timings(dpgp_handle* h, const char** names, float* ms, int cap) {
  if (!h) return 0;
  int k = 0;
  for (int i = 0; i < kNumPhases && k < cap; ++i) {
    if (!h->ev_used[i]) continue;
    float t = 0;
    if (cudaEventSynchronize(h->ev1[i]) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&t, h->ev0[i], h->ev1[i]) != cudaSuccess) continue;
    names[k] = kPhaseNames[i]; ms[k] = t; ++k;
  }
  return k;
}

// ------------------------------------------------------------------------------------------ kernel API
namespace {
__global__ void covariance_kernel(const double* x0, int64_t n0, const double* x1, int64_t n1, const double* gamma,
                                  const double* alpha, const double* beta, int q, int b_count, int noise, int jitter,
                                  double* out) {
  const int64_t total = (int64_t)b_count * n0 * n1;
  const bool square = (x1 == nullptr);
  const double* xb = square ? x0 : x1;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / (n0 * n1));
    const int64_t rem = idx % (n0 * n1), i = rem / n1, j = rem % n1;
    double xi = 0, xj = 0, xx = 0;
    for (int k = 0; k < q; ++k) {
      const double sg = sqrt(gamma[b * q + k]);
      const double a = sg * x0[i * q + k], c = sg * xb[j * q + k];
      xi = fma(a, a, xi); xj = fma(c, c, xj); xx = fma(a, c, xx);
    }
    double k = alpha[b] * exp(-0.5 * xi - 0.5 * xj + xx);
    if (square && i == j) { if (noise) k += 1.0 / beta[b]; if (jitter) k += kJitter; }
    out[idx] = k;
  }
}
}  // namespace

int dpgp_covariance(dpgp_handle* h, const double* d_x0, int64_t n0, const double* d_x1, int64_t n1,
                    const double* d_gamma, const double* d_alpha, const double* d_beta,
                    int include_noise, int include_jitter, double* d_out, void* stream) {
  if (!h || !d_x0 || !d_gamma || !d_alpha || !d_out || n0 < 1) return fail(h, DPGP_E_ARG, "dpgp_covariance: null/empty argument");
  if (!d_x1) n1 = n0;
  if (include_noise && !d_beta) return fail(h, DPGP_E_ARG, "dpgp_covariance: include_noise needs beta");
  const int64_t total = (int64_t)h->b * n0 * n1;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)h->sms * 16);
  covariance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x0, n0, d_x1, n1, d_gamma, d_alpha, d_beta, h->q, h->b,
                                                            include_noise, include_jitter, d_out);
  POST_LAUNCH(h, "covariance_kernel");
  return DPGP_OK;
}

namespace {
int launch_psi1_fwd(dpgp_handle* h, const double* mu, const double* s, const double* y, const double* z,
                    const double* gamma, const double* alpha, int64_t n, double* psi1_out, double* p_out, cudaStream_t st) {
  Psi1FwdParams p{};
  p.mu = mu; p.s = s; p.y = y; p.z = z; p.gamma = gamma; p.alpha = alpha;
  p.part = h->p1_part; p.tags = h->p1_tags; p.psi1_out = psi1_out;
  p.n = n; p.d = h->d; p.q = h->q; p.m = h->m; p.mp = h->mp; p.b = h->b; p.mode = h->mode; p.ncols = h->ncols; p.cpad = h->cpad;
  p.nchunks = cdiv64(n, kP1Rows); p.nseg = h->p1_nseg;
  const size_t smem = ((size_t)kP1Rows * h->mp + (size_t)kP1Cols * kP1Rows) * 8;
  h->k->psi1_fwd(h->p1_grid, smem, st, p);
  POST_LAUNCH(h, "psi1_fwd_kernel");
  if (p_out) {
    PReduceParams r{h->p1_part, h->p1_tags, p_out, h->p1_grid, h->p1_nseg, h->m, h->mp, h->ncols, h->cpad, h->b, p.nchunks};
    const int total = h->b * h->m * h->ncols;
    p_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(r);
    POST_LAUNCH(h, "p_reduce_kernel");
  }
  return DPGP_OK;
}
}  // namespace

int dpgp_psi1(dpgp_handle* h, const double* d_mu, const double* d_s, int64_t n, const double* d_z,
              const double* d_gamma, const double* d_alpha, double* d_out, void* stream) {
  if (!h || !d_mu || !d_s || !d_z || !d_gamma || !d_alpha || !d_out || n < 1) return fail(h, DPGP_E_ARG, "dpgp_psi1: null/empty argument");
  // Y is not needed for the materialised statistic: contract against nothing by passing ncols = 0 columns
  Psi1FwdParams p{};
  p.mu = d_mu; p.s = d_s; p.y = nullptr; p.z = d_z; p.gamma = d_gamma; p.alpha = d_alpha;
  p.part = h->p1_part; p.tags = h->p1_tags; p.psi1_out = d_out;
  p.n = n; p.d = h->d; p.q = h->q; p.m = h->m; p.mp = h->mp; p.b = h->b; p.mode = h->mode; p.ncols = 0; p.cpad = h->cpad;
  if (n != h->n) return fail(h, DPGP_E_ARG, "dpgp_psi1: n (%lld) must equal the handle's n_local (%lld)", (long long)n, (long long)h->n);
  p.nchunks = cdiv64(n, kP1Rows); p.nseg = h->p1_nseg;
  const size_t smem = ((size_t)kP1Rows * h->mp + (size_t)kP1Cols * kP1Rows) * 8;
  h->k->psi1_fwd(h->p1_grid, smem, (cudaStream_t)stream, p);
  POST_LAUNCH(h, "psi1_fwd_kernel");
  return DPGP_OK;
}

// ------------------------------------------------------------------------------------------- hot path
int dpgp_stats_fwd(dpgp_handle* h, const double* d_mu, const double* d_s, const double* d_y, const double* d_z,
                   const double* d_gamma, const double* d_alpha, double* d_stats, void* stream) {
  if (!h || !d_mu || !d_s || !d_y || !d_z || !d_gamma || !d_alpha || !d_stats) return fail(h, DPGP_E_ARG, "dpgp_stats_fwd: null argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->fine) { h->fine_stream = st; fine_mark(h, "(start of dpgp_stats_fwd)"); }
  if (h->side) {
    // fork: the K_uu factor needs none of the statistics; it runs next to prep / psi2 forward and is joined by dpgp_bound
    CU(h, cudaEventRecord(h->ev_fork, st));
    CU(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    if (int rc = kuu_factor(h, d_z, d_gamma, d_alpha, h->side)) return rc;
    CU(h, cudaEventRecord(h->ev_kuu, h->side));
    h->kuu_pending = true; h->kuu_z = d_z; h->kuu_gamma = d_gamma; h->kuu_alpha = d_alpha;
  }
  double* psi2 = d_stats;
  double* pm = psi2 + (size_t)h->b * h->m * h->m;
  double* yy = pm + (size_t)h->b * h->m * h->ncols;
  {
    PhaseTimer t(h, PH_PREP, st);
    PrepParams p{d_mu, d_s, d_z, d_gamma, d_alpha, h->r, h->v, h->n, h->q, h->m, h->mp, h->b, cdiv64(h->n, kPrepRows)};
    const int grid = (int)std::min<int64_t>(p.nchunks * h->b, (int64_t)h->sms * 8);
    h->k->prep(grid, st, p);
    POST_LAUNCH(h, "prep_rows_kernel");
  }
  {
    PhaseTimer t(h, PH_PSI2F, st);
    Psi2FwdParams p{};
    p.r = h->r; p.v = h->v; p.z = d_z; p.part = h->f_part; p.tags = h->f_tags; p.exptab = h->expv_fwd == 8 ? h->exptab_fwd : h->exptab;
    p.n = h->n; p.q = h->q; p.m = h->m; p.mp = h->mp; p.mt = h->mt; p.b = h->b; p.npass = h->f_npass;
    p.chunk = h->f_chunk; p.nchunks = cdiv64(h->n, h->f_chunk); p.nseg = h->f_nseg;
    for (int sp = 0; sp < h->f_splits; ++sp) {
      p.tile0 = sp * h->f_t2; p.t2 = std::min(h->f_t2, h->t2 - p.tile0);
      if (p.t2 <= 0) break;
      h->k->psi2_fwd(h->expv_fwd, h->grid, h->f_threads, h->f_smem, st, p);
      POST_LAUNCH(h, "psi2_fwd_kernel");
      Psi2ReduceParams r{h->f_part, h->f_tags, psi2, h->grid, h->f_nseg, h->f_npass * (h->f_threads - 32) * 4, h->m, h->mt, p.t2, h->b, p.tile0, p.nchunks};
      const int total = h->b * p.t2 * 4;
      psi2_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(r);
      POST_LAUNCH(h, "psi2_reduce_kernel");
    }
  }
  {
    PhaseTimer t(h, PH_PSI1F, st);
    int rc = launch_psi1_fwd(h, d_mu, d_s, d_y, d_z, d_gamma, d_alpha, h->n, nullptr, pm, st);
    if (rc) return rc;
    ColSumParams c{d_mu, d_s, d_y, h->cs_part, h->n, h->d, h->q};
    colsum_kernel<<<h->cs_grid, 256, 0, st>>>(c);
    POST_LAUNCH(h, "colsum_kernel");
    colsum_reduce_kernel<<<(h->d + 2 + 255) / 256, 256, 0, st>>>(h->cs_part, yy, h->cs_grid, h->d + 2);
    POST_LAUNCH(h, "colsum_reduce_kernel");
  }
  return DPGP_OK;
}

int dpgp_bound(dpgp_handle* h, int64_t n_total, const double* d_stats, const double* d_z, const double* d_gamma,
               const double* d_alpha, const double* d_beta, const double* d_wgt, double* d_gp, double* d_dstats,
               double* d_dz, double* d_dgamma, double* d_dalpha, double* d_dbeta, double* d_dwgt, void* stream) {
  if (!h || !d_stats || !d_z || !d_gamma || !d_alpha || !d_beta || !d_gp || !d_dstats || !d_dz || !d_dgamma || !d_dalpha || !d_dbeta)
    return fail(h, DPGP_E_ARG, "dpgp_bound: null argument");
  if (h->mode == DPGP_MODE_T && (!d_wgt || !d_dwgt)) return fail(h, DPGP_E_ARG, "dpgp_bound: T-mode needs phi and its gradient buffer");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->fine) { h->fine_stream = st; fine_mark(h, "(start of dpgp_bound)"); }
  PhaseTimer t(h, PH_BOUND, st);
  const size_t mm = (size_t)h->m * h->m, mc = (size_t)h->m * h->ncols;
  const double* psi2 = d_stats; const double* pm = psi2 + h->b * mm; const double* yy = pm + h->b * mc; const double* kl = yy + h->d;
  double* dpsi2 = d_dstats; double* dp = dpsi2 + h->b * mm; double* dyy = dp + h->b * mc; double* dkl = dyy + h->d;
  const double* wgt = (h->mode == DPGP_MODE_T) ? d_wgt : nullptr;
  const int M = h->m, C = h->ncols;
  const long long lmm = (long long)mm, lmc = (long long)mc;
  if (h->kuu_pending && h->kuu_z == d_z && h->kuu_gamma == d_gamma && h->kuu_alpha == d_alpha) {
    CU(h, cudaStreamWaitEvent(st, h->ev_kuu, 0));            // join the side stream (dpgp_stats_fwd started the factor there)
  } else {
    if (h->kuu_pending) CU(h, cudaStreamWaitEvent(st, h->ev_kuu, 0));      // different point: join, then redo in line
    if (int rc = kuu_factor(h, d_z, d_gamma, d_alpha, st)) return rc;
  }
  h->kuu_pending = false;
  int rc;
  // T1 = L^-1 Psi2, C1 = L^-1 P;  H^T = L^-1 T1^T  (the reference's two triangular solves, dp_gp_lvm.py:118-121 / :621-624)
  if ((rc = launch_trsm(h, {ts_job(h->lk, lmm, psi2, lmm, M, 1, h->t1, lmm, M, M), ts_job(h->lk, lmm, pm, lmc, C, 1, h->c1, lmc, C, C)}, st))) return rc;
  if ((rc = launch_trsm(h, {ts_job(h->lk, lmm, h->t1, lmm, 1, M, h->hmat, lmm, M, M)}, st))) return rc;
  // L_A = chol(beta H + I), Lainv = L_A^-1, log det L_A, tr H, tr A^-1; mirrors the lower triangle of H (the one it factors)
  // into the upper one, so that the cotangents below differentiate exactly the function that was evaluated
  {
    FactorParams f{};
    f.hmat = h->hmat; f.beta = d_beta; f.out = h->lainv; f.lout = h->la; f.scal = h->fscal; f.bad = h->bad; f.mode = 1; f.m = M; f.q = h->q; f.bad_offset = 1000;
    if ((rc = launch_factor(h, f, st))) return rc;
  }
  // Cm = L_A^-1 C1 (dp_gp_lvm.py:132-133 / :638-639)
  if ((rc = launch_trsm(h, {ts_job(h->la, lmm, h->c1, lmc, C, 1, h->cm, lmc, C, C)}, st))) return rc;
  // G2 = Lainv H;  R = Lainv Linv
  if ((rc = launch_jobs(h, {mm_job(h->lainv, lmm, M, 1, h->hmat, lmm, M, 1, h->t1, lmm, M, M, M, M, MM_A_LOWER),
                            mm_job(h->lainv, lmm, M, 1, h->linv, lmm, M, 1, h->rmat, lmm, M, M, M, M, MM_A_LOWER | MM_B_KGEJ)}, st))) return rc;
  // G = G2 Linv = Lainv H Linv;  U = R^T Cm = S P
  if ((rc = launch_jobs(h, {mm_job(h->t1, lmm, M, 1, h->linv, lmm, M, 1, h->gmat, lmm, M, M, M, M, MM_B_KGEJ),
                            mm_job(h->rmat, lmm, 1, M, h->cm, lmc, C, 1, h->umat, lmc, C, M, C, M, MM_A_UPPER)}, st))) return rc;
  // R^T G = (Kinv - S) / beta;  G^T G = -(Kinv - S - beta Kinv Psi2 Kinv) / beta^2;  PU = Psi2 U;  W = U diag(w) U^T
  if ((rc = launch_jobs(h, {mm_job(h->rmat, lmm, 1, M, h->gmat, lmm, M, 1, h->rtg, lmm, M, M, M, M, MM_A_UPPER),
                            mm_job(h->gmat, lmm, 1, M, h->gmat, lmm, M, 1, h->gtg, lmm, M, M, M, M, MM_SYM),
                            mm_job(psi2, lmm, M, 1, h->umat, lmc, C, 1, h->pu, lmc, C, M, C, M, 0),
                            mm_job(h->umat, lmc, C, 1, h->umat, lmc, 1, C, h->wmat, lmm, M, M, M, C, MM_SYM, wgt, 1, h->b)}, st))) return rc;
  {
    BoundOutParams o{};
    o.rtg = h->rtg; o.gtg = h->gtg; o.wmat = h->wmat; o.g2 = h->t1; o.lainv = h->lainv; o.cm = h->cm; o.u = h->umat; o.pu = h->pu; o.scal = h->fscal;
    o.yy = yy; o.alpha = d_alpha; o.beta = d_beta; o.wgt = wgt;
    o.fb = h->fb; o.dpsi2 = dpsi2; o.dp = dp; o.dk = h->dk; o.dbeta = d_dbeta; o.dalpha_direct = h->dadirect;
    o.dwgt = (h->mode == DPGP_MODE_T) ? d_dwgt : nullptr;
    o.n_total = n_total; o.d = h->d; o.m = M; o.b = h->b; o.mode = h->mode; o.ncols = C;
    const int eb = (int)std::min<size_t>(16, (std::max(mm, mc) + 1023) / 1024);
    bound_out_kernel<<<dim3(1 + eb, h->b), 256, 0, st>>>(o);
    POST_LAUNCH(h, "bound_out_kernel");
  }
  BoundFinishParams f{h->fb, kl, d_beta, wgt, d_gp, dyy, dkl, n_total, h->d, h->q, h->b, h->mode};
  bound_finish_kernel<<<1, 256, 0, st>>>(f);
  POST_LAUNCH(h, "bound_finish_kernel");
  ZChainParams zc{h->dk, nullptr, d_z, d_gamma, d_alpha, h->dzk, h->zpart, h->q, h->qp, h->m, h->b};
  launch_zchain(h, zc, st);
  POST_LAUNCH(h, "zchain_kernel");
  // dz = sum_b dzk[b];  dgamma, dalpha from the row-block partials (+ direct term of dalpha)
  bound_fin_kernel<<<(h->m * h->q + h->b * (h->q + 1) + 255) / 256, 256, 0, st>>>(h->dzk, h->dadirect, h->zpart, d_alpha, d_dz, d_dgamma,
                                                                                  d_dalpha, h->b, h->m * h->q, h->q, (h->m + kZcRows - 1) / kZcRows);
  POST_LAUNCH(h, "bound_fin_kernel");
  return DPGP_OK;
}

namespace {
__global__ void bound_factors_kernel(const double* kin, const double* sin_, const double* uin, double* kinv, double* sinv, double* u,
                                     size_t nmm, size_t nmc) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nmm + nmc; i += (size_t)gridDim.x * blockDim.x) {
    if (i < nmm) { if (kinv) kinv[i] = kin[i]; if (sinv) sinv[i] = sin_[i]; }
    else if (u) u[i - nmm] = uin[i - nmm];
  }
}
}  // namespace

int dpgp_bound_factors(dpgp_handle* h, double* d_kinv, double* d_sinv, double* d_u, void* stream) {
  if (!h) return DPGP_E_ARG;
  const size_t nmm = (size_t)h->b * h->m * h->m, nmc = (size_t)h->b * h->m * h->ncols;
  {
    // Kinv = Linv^T Linv and S = R^T R are not needed by the bound itself any more (its cotangents use the cancellation-free
    // forms R^T G and G^T G); they are formed here, from the factors of the most recent dpgp_bound call.
    const int M = h->m; const long long lmm = (long long)M * M;
    if (int rc = launch_jobs(h, {mm_job(h->linv, lmm, 1, M, h->linv, lmm, M, 1, h->kinv, lmm, M, M, M, M, MM_A_UPPER | MM_B_KGEJ | MM_SYM),
                                 mm_job(h->rmat, lmm, 1, M, h->rmat, lmm, M, 1, h->smat, lmm, M, M, M, M, MM_A_UPPER | MM_B_KGEJ | MM_SYM)},
                             (cudaStream_t)stream)) return rc;
  }
  bound_factors_kernel<<<std::min(h->sms * 4, 1024), 256, 0, (cudaStream_t)stream>>>(h->kinv, h->smat, h->umat, d_kinv, d_sinv, d_u, nmm, nmc);
  POST_LAUNCH(h, "bound_factors_kernel");
  return DPGP_OK;
}

int dpgp_stats_bwd(dpgp_handle* h, const double* d_mu, const double* d_s, const double* d_y, const double* d_z,
                   const double* d_gamma, const double* d_alpha, const double* d_dstats, double* d_dmu, double* d_ds,
                   double* d_dz, double* d_dgamma, double* d_dalpha, void* stream) {
  if (!h || !d_mu || !d_s || !d_y || !d_z || !d_gamma || !d_alpha || !d_dstats || !d_dmu || !d_ds || !d_dz || !d_dgamma || !d_dalpha)
    return fail(h, DPGP_E_ARG, "dpgp_stats_bwd: null argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->fine) { h->fine_stream = st; fine_mark(h, "(start of dpgp_stats_bwd)"); }
  const size_t mm = (size_t)h->m * h->m, mc = (size_t)h->m * h->ncols;
  const double* dpsi2 = d_dstats; const double* dp = dpsi2 + h->b * mm; const double* dkl = dp + h->b * mc + h->d;
  // r / v must be those of the same parameter point: dpgp_stats_fwd of this evaluation produced them.
  if (h->bwd_variant != 2) {
    PhaseTimer t(h, PH_BWDF, st);
    Psi2BwdFusedParams p{};
    p.r = h->r; p.v = h->v; p.z = d_z; p.gbar = dpsi2; p.exptab = h->exptab; p.sched = h->u_sched;
    p.dr = h->r /* in place */; p.dv = h->dv; p.part = h->u_part; p.tags = h->u_tags;
    p.n = h->n; p.q = h->q; p.m = h->m; p.mp = h->mp; p.b = h->b; p.nrounds = h->u_nrounds; p.nseg = h->u_nseg;
    p.ngroups = cdiv64(h->n, 32 * h->u_rows);
    CU(h, cudaMemsetAsync(h->u_part, 0, sizeof(double) * h->u_grid * h->u_nseg * h->u_slice, st));
#ifdef DPGP_EXPERIMENTAL
    if (h->bwd_variant == 5) h->k->psi2_bwd_ws(h->expv, h->u_grid, h->u_smem, st, p, false);
    else if (h->bwd_variant == 4) h->k->psi2_bwd_fused2(h->expv, h->u_grid, h->u_smem, st, p, false);
    else if (h->bwd_variant == 3) h->k->psi2_bwd_tc(h->expv, h->u_grid, h->u_smem, st, p, false);
    else
#endif
    if (h->bwd_variant == 7) {
      UmmaTablesParams tp{dpsi2, d_z, h->u_sched, h->um_wtab, h->um_dprime, h->um_scale, h->q, h->m, h->u_nrounds};
      umma_tables_kernel<<<h->b, 256, 0, st>>>(tp);
      POST_LAUNCH(h, "umma_tables_kernel");
      Psi2BwdUmmaParams up{p, h->um_wtab, h->um_dprime, h->um_scale, nullptr, getenv("DPGP_UM_SKIP") ? atoi(getenv("DPGP_UM_SKIP")) : 0};
      static long long* um_prof = nullptr;                 // development: DPGP_UM_PROF=1 prints the role counters of CTA 0 after each launch
      if (getenv("DPGP_UM_PROF")) { if (!um_prof) cudaMalloc(&um_prof, 32 * sizeof(long long)); cudaMemsetAsync(um_prof, 0, 32 * sizeof(long long), st); up.prof = um_prof; }
      h->k->psi2_bwd_umma(h->expv, h->u_grid, h->um_smem, st, up, false);
      POST_LAUNCH(h, "psi2_bwd_umma_kernel");
      if (up.prof) {
        long long hp[32]; cudaStreamSynchronize(st); cudaMemcpy(hp, up.prof, sizeof hp, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[um prof, CTA 0, cycles] producers total/wait-done:");
        for (int w = 0; w < 8; ++w) fprintf(stderr, " %lld/%lld", hp[2 * w], hp[2 * w + 1]);
        fprintf(stderr, "\n  mma total %lld wait-full %lld wait-ddempty %lld | drain0 total %lld wait-done %lld read %lld | drain1 total %lld wait-done %lld read %lld\n",
                hp[16], hp[17], hp[18], hp[20], hp[21], hp[22], hp[23], hp[24], hp[25]);
      }
    } else if (h->bwd_variant == 8) {
      h->k->psi2_bwd_mma(h->expv, h->u_grid, h->u_smem, st, p, false);
      POST_LAUNCH(h, "psi2_bwd_mma_kernel");
    } else {
    h->k->psi2_bwd_fused(h->expv, h->u_rows, h->u_grid, h->u_smem, st, p, h->bwd_variant == 6);
    POST_LAUNCH(h, "psi2_bwd_fused_kernel");
    }
    if (h->bwd_variant >= 6) {
      DzFusedReduceParams r{h->u_part, h->dzd, h->u_grid * kFusedWarps, h->m, h->mp, h->q, h->qp, h->b};
      const int64_t warps = (int64_t)h->b * h->m * h->q;
      dz_fused_reduce_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(r);
      POST_LAUNCH(h, "dz_fused_reduce_kernel");
    } else {
    DdFusedReduceParams r{h->u_part, h->u_tags, h->u_sched, h->ddsym, h->u_grid, h->u_nseg, h->u_nrounds, h->m, h->b, h->qp, p.ngroups};
    const int64_t total = (int64_t)h->b * (int64_t)h->u_slice;
    CU(h, cudaMemsetAsync(h->ddsym, 0, sizeof(double) * h->b * mm * h->qp, st));
    dd_fused_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(r);
    POST_LAUNCH(h, "dd_fused_reduce_kernel");
    }
  } else {
#ifdef DPGP_EXPERIMENTAL
  {
    PhaseTimer t(h, PH_BWDP, st);
    Psi2BwdPairParams p{};
    p.r = h->r; p.v = h->v; p.z = d_z; p.gbar = dpsi2; p.part = h->bp_part; p.tags = h->bp_tags; p.exptab = h->exptab;
    p.n = h->n; p.q = h->q; p.m = h->m; p.mp = h->mp; p.mt = h->mt; p.b = h->b; p.t2 = h->t2; p.jb = h->p_jb; p.ng = h->p_ng;
    p.chunk = h->p_chunk; p.nchunks = cdiv64(h->n, h->p_chunk); p.nseg = h->p_nseg;
    const int pgrid = h->p_jb * h->p_ng;
    h->k->psi2_bwd_pair(h->expv, pgrid, h->p_threads, h->p_smem, st, p);
    POST_LAUNCH(h, "psi2_bwd_pair_kernel");
    DdReduceParams r{h->bp_part, h->bp_tags, h->ddsym, pgrid, h->p_jb, h->p_threads - 32, h->m, h->mt, h->t2, h->b, h->qp, h->p_nseg};
    const int64_t total = (int64_t)h->b * 2 * h->t2 * 2 * h->qp;
    CU(h, cudaMemsetAsync(h->ddsym, 0, sizeof(double) * h->b * mm * h->qp, st));
    dd_reduce_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(r);
    POST_LAUNCH(h, "dd_reduce_kernel");
  }
  {
    PhaseTimer t(h, PH_BWDN, st);
    BlockTabParams bt{d_z, dpsi2, h->dtab, h->gtab, h->q, h->qp, h->m, h->mp, h->b};
    block_tables_kernel<<<h->sms, 256, 0, st>>>(bt);
    POST_LAUNCH(h, "block_tables_kernel");
    Psi2BwdNParams p{};
    p.r = h->r; p.v = h->v; p.dtab = h->dtab; p.gtab = h->gtab; p.dr = h->r /* in place */; p.dv = h->dv; p.exptab = h->exptab;
    p.n = h->n; p.q = h->q; p.m = h->m; p.mp = h->mp; p.b = h->b; p.ngroups = cdiv64(h->n, h->n_threads);
    const int grid = (int)std::min<int64_t>(p.ngroups * h->b, (int64_t)h->grid);
    h->k->psi2_bwd_n(h->expv, grid, h->n_threads, h->n_smem, st, p);
    POST_LAUNCH(h, "psi2_bwd_n_kernel");
  }
#endif
  }
  {
    PhaseTimer t(h, PH_CHAIN, st);
    int cgrid = 0;
    if (h->chain_variant == 1) {
      Chain2Params c{};
      c.mu = d_mu; c.s = d_s; c.y = d_y; c.z = d_z; c.gamma = d_gamma; c.alpha = d_alpha; c.dp = dp; c.dr = h->r; c.dv = h->dv; c.dkl = dkl;
      c.dmu = d_dmu; c.ds = d_ds; c.dzp = h->dzp; c.dgp = h->dgp; c.dap = h->dap;
      c.n = h->n; c.d = h->d; c.q = h->q; c.m = h->m; c.mp = h->mp; c.b = h->b; c.mode = h->mode; c.ncols = h->ncols;
      c.nchunks = cdiv64(h->n, h->c2_rows);
      c.bgroups = h->c2_groups; c.cgrid = h->c2_grid; c.dmu_part = h->c2_mu_part; c.ds_part = h->c2_s_part;
      cgrid = h->c2_grid * h->c2_groups;
      if (h->c2_groups > 1) {                       // a CTA only writes the partials of its own clusters
        CU(h, cudaMemsetAsync(h->dgp, 0, sizeof(double) * (size_t)cgrid * h->b * h->qp, st));
        CU(h, cudaMemsetAsync(h->dap, 0, sizeof(double) * (size_t)cgrid * h->b, st));
      }
      h->k->chain2(h->c2_rows, cgrid, h->c2_smem, st, c);
      POST_LAUNCH(h, "psi1_bwd_chain_kernel");
      if (h->c2_groups > 1) {
        const int64_t len = h->n * h->q;
        chain2_rows_reduce_kernel<<<(int)std::min<int64_t>((len + 255) / 256, 1024), 256, 0, st>>>(h->c2_mu_part, h->c2_s_part, d_dmu, d_ds, len, h->c2_groups);
        POST_LAUNCH(h, "chain2_rows_reduce_kernel");
      }
    } else {
#ifdef DPGP_EXPERIMENTAL
    G1Params g{};
    g.mu = d_mu; g.s = d_s; g.y = d_y; g.z = d_z; g.gamma = d_gamma; g.alpha = d_alpha; g.dp = dp; g.bco = h->bco;
    g.n = h->n; g.d = h->d; g.q = h->q; g.m = h->m; g.mp = h->mp; g.b = h->b; g.mode = h->mode; g.ncols = h->ncols;
    g.nchunks = cdiv64(h->n, kP1Rows);
    const size_t g1_smem = ((size_t)kP1Rows * h->mp + (size_t)kP1Cols * kP1Rows + (size_t)kP1Cols * h->mp) * 8;
    const int ggrid = (int)std::min<int64_t>(g.nchunks * h->b, (int64_t)h->sms * 4);
    h->k->g1(ggrid, g1_smem, st, g);
    POST_LAUNCH(h, "g1_kernel");
    ChainParams c{};
    c.mu = d_mu; c.s = d_s; c.z = d_z; c.gamma = d_gamma; c.alpha = d_alpha; c.dr = h->r; c.dv = h->dv; c.bco = h->bco; c.dkl = dkl;
    c.dmu = d_dmu; c.ds = d_ds; c.dzp = h->dzp; c.dgp = h->dgp; c.dap = h->dap;
    c.n = h->n; c.q = h->q; c.m = h->m; c.mp = h->mp; c.b = h->b; c.nchunks = cdiv64(h->n, kChRows);
    const size_t ch_smem = (2 * (size_t)kChRows * h->mp + (size_t)h->mp * h->qp) * 8;
    cgrid = (int)std::min<int64_t>(c.nchunks, (int64_t)h->grid);
    h->k->chain(cgrid, ch_smem, st, c);
    POST_LAUNCH(h, "chain_bwd_kernel");
#endif
    }
    if (h->bwd_variant < 6) {                           // variants 6 / 7 have filled dzd already (dz_fused_reduce_kernel)
      ZChainParams zc{nullptr, h->ddsym, d_z, d_gamma, d_alpha, h->dzd, nullptr, h->q, h->qp, h->m, h->b};
      launch_zchain(h, zc, st);
      POST_LAUNCH(h, "zchain_kernel");
    }
    PhaseTimer t2(h, PH_REDUCE, st);
    const int total = h->m * h->q + h->b * h->q + h->b;
    if (h->chain_variant == 1)
      chain2_reduce_kernel<<<(total * 32 + 255) / 256, 256, 0, st>>>(h->dzp, h->dgp, h->dap, h->dzd, d_dz, d_dgamma, d_dalpha, cgrid,
                                                                 h->b, h->m, h->mp, h->q, h->qp);
    else
      chain_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(h->dzp, h->dgp, h->dap, h->dzd, d_dz, d_dgamma, d_dalpha, cgrid,
                                                                h->b, h->m, h->mp, h->q, h->qp);
    POST_LAUNCH(h, "chain_reduce_kernel");
  }
  return DPGP_OK;
}

}  // extern "C"
