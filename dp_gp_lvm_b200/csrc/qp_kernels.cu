// Compiled once per -DDPGP_QP=<padded latent dimension>; see qp_kernels.cuh.
#ifndef DPGP_QP
#error "compile with -DDPGP_QP=<2|4|6|8|10|12|16|20|24|28|32>"
#endif
#include <cstdlib>

#include "qp_kernels.cuh"

#define DPGP_CAT_(a, b) a##b
#define DPGP_CAT(a, b) DPGP_CAT_(a, b)

namespace dpgp {
namespace {
constexpr int QP = DPGP_QP;

#ifdef DPGP_EXPERIMENTAL
#define EXP_SWITCH(EV, ...)                                   \
  switch (EV) {                                               \
    case 1: { constexpr int EXPV = 1; __VA_ARGS__; break; }   \
    case 3: { constexpr int EXPV = 3; __VA_ARGS__; break; }   \
    case 2: { constexpr int EXPV = 2; __VA_ARGS__; break; }   \
    case 5: { constexpr int EXPV = 5; __VA_ARGS__; break; }   \
    case 6: { constexpr int EXPV = 6; __VA_ARGS__; break; }   \
    default: { constexpr int EXPV = 4; __VA_ARGS__; break; }  \
  }
#else
#define EXP_SWITCH(EV, ...)                                   \
  switch (EV) {                                               \
    case 1: { constexpr int EXPV = 1; __VA_ARGS__; break; }   \
    default: { constexpr int EXPV = 4; __VA_ARGS__; break; }  \
  }
#endif

template <typename K>
cudaError_t optin(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// 64-row groups (two rows per lane) in the fused backward: registers allow it up to QP = 12.  The helpers below are
// templates only so that `if constexpr` discards the R = 2 instantiations for larger QP.
template <int Q_> constexpr bool kTwoRows = Q_ <= 12;

template <int Q_ = QP>
size_t fused_smem_t(int rows, int mp) {
  if constexpr (kTwoRows<Q_>) { if (rows == 2) return fused_smem_bytes<Q_, 2>(mp); }
  return fused_smem_bytes<Q_, 1>(mp);
}
size_t fused_smem(int rows, int mp) { return fused_smem_t<>(rows, mp); }
template <int Q_ = QP>
cudaError_t cfg_smem_t(int expv, size_t f, size_t p1, int urows, size_t fused) {
  cudaError_t e;
  if ((e = optin(psi1_fwd_kernel<QP, true>, p1)) != cudaSuccess) return e;
  if ((e = optin(psi1_fwd_tc_kernel<QP>, psi1_tc_smem_bytes(128))) != cudaSuccess) return e;
  if ((e = optin(psi1_fwd_kernel<QP, false>, p1)) != cudaSuccess) return e;
  EXP_SWITCH(expv, {
    if ((e = optin(psi2_fwd_kernel<QP, EXPV>, f)) != cudaSuccess) return e;
    if ((e = optin(psi2_fwd_kernel<QP, 8>, f)) != cudaSuccess) return e;
    if constexpr (kTwoRows<Q_>) {
      if (urows == 2) {
        if ((e = optin(psi2_bwd_fused_kernel<QP, EXPV, 2>, fused)) != cudaSuccess) return e;
        if ((e = optin(psi2_bwd_fused_kernel<QP, EXPV, 2, 1, kFusedKuDz, true>, fused)) != cudaSuccess) return e;
      }
    }
    if (urows != 2) {
      if ((e = optin(psi2_bwd_fused_kernel<QP, EXPV, 1>, fused)) != cudaSuccess) return e;
      if ((e = optin(psi2_bwd_fused_kernel<QP, EXPV, 1, 1, kFusedKuDz, true>, fused)) != cudaSuccess) return e;
    }
  });
  return cudaSuccess;
}
cudaError_t cfg_smem(int expv, size_t f, size_t p1, int urows, size_t fused) { return cfg_smem_t<>(expv, f, p1, urows, fused); }
void run_prep(int grid, cudaStream_t st, const PrepParams& p) { prep_rows_kernel<QP><<<grid, 256, 0, st>>>(p); }
void run_psi2_fwd(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2FwdParams& p) {
  if (expv == 8) { psi2_fwd_kernel<QP, 8><<<grid, threads, smem, st>>>(p); return; }      // 2 048-entry table (forward only)
  EXP_SWITCH(expv, { psi2_fwd_kernel<QP, EXPV><<<grid, threads, smem, st>>>(p); });
}
template <int Q_ = QP>
void run_psi2_bwd_fused_t(int expv, int rows, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool dz) {
  EXP_SWITCH(expv, {
    if constexpr (kTwoRows<Q_>) {
      if (rows == 2) {
        if (dz) psi2_bwd_fused_kernel<QP, EXPV, 2, 1, kFusedKuDz, true><<<grid, kFusedWarps * 32, smem, st>>>(p);
        else psi2_bwd_fused_kernel<QP, EXPV, 2><<<grid, kFusedWarps * 32, smem, st>>>(p);
        break;
      }
    }
    if (dz) psi2_bwd_fused_kernel<QP, EXPV, 1, 1, kFusedKuDz, true><<<grid, kFusedWarps * 32, smem, st>>>(p);
    else psi2_bwd_fused_kernel<QP, EXPV, 1><<<grid, kFusedWarps * 32, smem, st>>>(p);
  });
}
void run_psi2_bwd_fused(int expv, int rows, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool dz) {
  run_psi2_bwd_fused_t<>(expv, rows, grid, smem, st, p, dz);
}
template <int Q_ = QP>
bool run_psi2_bwd_umma_t(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdUmmaParams& p, bool configure_only) {
  // the opt-in variants are instantiated for the default exp (shared-memory table) only
  if constexpr (Q_ <= 16) {
    if (expv != 4) return false;
    if (configure_only) return optin(psi2_bwd_umma_kernel<QP, 4>, smem) == cudaSuccess;
    psi2_bwd_umma_kernel<QP, 4><<<grid, kUmThreads, smem, st>>>(p);
    return true;
  } else {
    return false;
  }
}
bool run_psi2_bwd_umma(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdUmmaParams& p, bool configure_only) {
  return run_psi2_bwd_umma_t<>(expv, grid, smem, st, p, configure_only);
}
template <int Q_ = QP>
size_t mma_smem_t(int mp) {
  if constexpr (Q_ >= 8 && Q_ <= 12) return mma_smem_bytes<Q_>(mp);
  else return 0;
}
size_t mma_smem(int mp) { return mma_smem_t<>(mp); }
template <int Q_ = QP>
bool run_psi2_bwd_mma_t(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only) {
  if constexpr (Q_ >= 8 && Q_ <= 12) {
    if (expv != 4) return false;
    if (configure_only) return optin(psi2_bwd_mma_kernel<QP, 4>, smem) == cudaSuccess;
    psi2_bwd_mma_kernel<QP, 4><<<grid, kFusedWarps * 32, smem, st>>>(p);
    return true;
  } else {
    return false;
  }
}
bool run_psi2_bwd_mma(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only) {
  return run_psi2_bwd_mma_t<>(expv, grid, smem, st, p, configure_only);
}
void run_psi1_fwd(int grid, size_t smem, cudaStream_t st, const Psi1FwdParams& p) {
  const bool persist = p.ncols <= kP1Cols && (p.mp / 4) * (kP1Cols / 4) <= 2 * 256;
  if (persist && p.ncols >= 8 && p.mp <= 128 && !getenv("DPGP_NO_PSI1_TC")) {      // contraction on the FP64 tensor cores
    psi1_fwd_tc_kernel<QP><<<grid, 256, psi1_tc_smem_bytes(p.mp), st>>>(p);
    return;
  }
  if (persist) psi1_fwd_kernel<QP, true><<<grid, 256, smem, st>>>(p);
  else psi1_fwd_kernel<QP, false><<<grid, 256, smem, st>>>(p);
}

size_t chain2_smem(int rows, int mp) { return rows == 32 ? chain2_smem_bytes<QP, 32>(mp) : chain2_smem_bytes<QP, 16>(mp); }
cudaError_t chain2_cfg(int rows, size_t smem) {
  return rows == 32 ? optin(psi1_bwd_chain_kernel<QP, 32>, smem) : optin(psi1_bwd_chain_kernel<QP, 16>, smem);
}
void run_chain2(int rows, int grid, size_t smem, cudaStream_t st, const Chain2Params& p) {
  if (rows == 32) psi1_bwd_chain_kernel<QP, 32><<<grid, 256, smem, st>>>(p);
  else psi1_bwd_chain_kernel<QP, 16><<<grid, 256, smem, st>>>(p);
}

#ifdef DPGP_EXPERIMENTAL
cudaError_t cfg_smem_x(int expv, size_t pp, size_t nn, size_t g1, size_t ch) {
  cudaError_t e;
  if ((e = optin(g1_kernel<QP>, g1)) != cudaSuccess) return e;
  if ((e = optin(chain_bwd_kernel<QP>, ch)) != cudaSuccess) return e;
  EXP_SWITCH(expv, {
    if ((e = optin(psi2_bwd_pair_kernel<QP, EXPV>, pp)) != cudaSuccess) return e;
    if ((e = optin(psi2_bwd_n_kernel<QP, EXPV>, nn)) != cudaSuccess) return e;
  });
  return cudaSuccess;
}
void run_psi2_bwd_pair(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2BwdPairParams& p) {
  EXP_SWITCH(expv, { psi2_bwd_pair_kernel<QP, EXPV><<<grid, threads, smem, st>>>(p); });
}
void run_psi2_bwd_n(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2BwdNParams& p) {
  EXP_SWITCH(expv, { psi2_bwd_n_kernel<QP, EXPV><<<grid, threads, smem, st>>>(p); });
}
size_t fused2_smem(int mp) { return fused_smem_bytes<QP, 1, 2>(mp); }
bool run_psi2_bwd_fused2(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only) {
  bool ok = true;
  EXP_SWITCH(expv, {
    if (configure_only) ok = optin(psi2_bwd_fused_kernel<QP, EXPV, 1, 2>, smem) == cudaSuccess;
    else psi2_bwd_fused_kernel<QP, EXPV, 1, 2><<<grid, kFusedWarps * 64, smem, st>>>(p);
  });
  return ok;
}
size_t ws_smem(int mp) { if constexpr (QP <= 12) return ws_smem_bytes<QP>(mp); else return ~(size_t)0; }
bool run_psi2_bwd_ws(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only) {
  if constexpr (QP <= 12) {
    bool ok = true;
    EXP_SWITCH(expv, {
      if (configure_only) ok = optin(psi2_bwd_ws_kernel<QP, EXPV>, smem) == cudaSuccess;
      else psi2_bwd_ws_kernel<QP, EXPV><<<grid, kFusedWarps * 64, smem, st>>>(p);
    });
    return ok;
  } else {
    return false;
  }
}
bool run_psi2_bwd_tc(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only) {
  if constexpr (QP <= 12) {
    bool ok = true;
    EXP_SWITCH(expv, {
      if (configure_only) ok = optin(psi2_bwd_tc_kernel<QP, EXPV>, smem) == cudaSuccess;
      else psi2_bwd_tc_kernel<QP, EXPV><<<grid, kFusedWarps * 32, smem, st>>>(p);
    });
    return ok;
  } else {
    return false;
  }
}
void run_g1(int grid, size_t smem, cudaStream_t st, const G1Params& p) { g1_kernel<QP><<<grid, 256, smem, st>>>(p); }
void run_chain(int grid, size_t smem, cudaStream_t st, const ChainParams& p) { chain_bwd_kernel<QP><<<grid, 256, smem, st>>>(p); }
#endif

const QpLaunchers kTable = {cfg_smem, fused_smem, run_psi2_bwd_fused, run_psi2_bwd_umma, mma_smem, run_psi2_bwd_mma, run_prep, run_psi2_fwd, run_psi1_fwd, chain2_smem, chain2_cfg, run_chain2
#ifdef DPGP_EXPERIMENTAL
                            , cfg_smem_x, fused2_smem, run_psi2_bwd_fused2, ws_smem, run_psi2_bwd_ws, run_psi2_bwd_tc, run_psi2_bwd_pair,
                            run_psi2_bwd_n, run_g1, run_chain
#endif
};
}  // namespace

const QpLaunchers* DPGP_CAT(qp_launchers_, DPGP_QP)() { return &kTable; }
}  // namespace dpgp
