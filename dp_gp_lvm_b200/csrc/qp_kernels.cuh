// Per-latent-dimension kernel launchers.  qp_kernels.cu is compiled once per padded latent dimension
// (-DDPGP_QP=2,4,...,32) so the instantiation sets build in parallel; dpgp_api.cu selects the table for the
// handle's QP.  The default build holds the default kernels (exp_variant 4 + libdevice, bwd_variant 6 + 1);
// `make EXPERIMENTAL=1` adds the variants under csrc/experimental/.
#pragma once
#include <cuda_runtime.h>

#include "psi1.cuh"
#include "chain2.cuh"
#include "psi2.cuh"
#include "psi2_bwd_fused.cuh"
#include "psi2_bwd_umma.cuh"
#include "psi2_bwd_mma.cuh"
#ifdef DPGP_EXPERIMENTAL      // `make EXPERIMENTAL=1`: the non-default variants measured in profiles/r01_*.md
#include "experimental/chain.cuh"
#include "experimental/psi2_bwd.cuh"
#include "experimental/psi2_bwd_tc.cuh"
#include "experimental/psi2_bwd_ws.cuh"
#endif

namespace dpgp {

struct QpLaunchers {
  // opt-in to > 48 KB of dynamic shared memory for every instantiation the handle can launch
  cudaError_t (*cfg_smem)(int expv, size_t f, size_t p1, int urows, size_t fused);
  size_t (*fused_smem)(int rows, int mp);
  void (*psi2_bwd_fused)(int expv, int rows, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool dz);
  // bwd_variant 7: dv / dD on the tcgen05 tensor cores as int8 slice products (QP <= 16, Mp <= 128); false if not instantiated
  bool (*psi2_bwd_umma)(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdUmmaParams& p, bool configure_only);
  // bwd_variant 8: dv / dD contractions as FP64 DMMA on q < 8 (8 <= QP <= 12); returns 0 if not instantiated for this QP
  size_t (*mma_smem)(int mp);
  bool (*psi2_bwd_mma)(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only);
  void (*prep)(int grid, cudaStream_t st, const PrepParams& p);
  void (*psi2_fwd)(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2FwdParams& p);
  void (*psi1_fwd)(int grid, size_t smem, cudaStream_t st, const Psi1FwdParams& p);
  size_t (*chain2_smem)(int rows, int mp);
  cudaError_t (*chain2_cfg)(int rows, size_t smem);
  void (*chain2)(int rows, int grid, size_t smem, cudaStream_t st, const Chain2Params& p);
#ifdef DPGP_EXPERIMENTAL
  cudaError_t (*cfg_smem_x)(int expv, size_t pp, size_t nn, size_t g1, size_t ch);
  // two teams of 8 warps per CTA on 32-row groups (16 warps / SM); configure_only sets the shared-memory attribute
  size_t (*fused2_smem)(int mp);
  bool (*psi2_bwd_fused2)(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only);
  // warp-specialised: 8 producer warps (phase 1) + 8 helper warps (phase 2), setmaxnreg
  size_t (*ws_smem)(int mp);
  bool (*psi2_bwd_ws)(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only);
  // tensor-core formulation of the first phase (QP <= 12, 64-row groups); returns false if not instantiated for this QP
  bool (*psi2_bwd_tc)(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only);
  void (*psi2_bwd_pair)(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2BwdPairParams& p);
  void (*psi2_bwd_n)(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2BwdNParams& p);
  void (*g1)(int grid, size_t smem, cudaStream_t st, const G1Params& p);
  void (*chain)(int grid, size_t smem, cudaStream_t st, const ChainParams& p);
#endif
};

// one table per padded latent dimension the Makefile builds (QPS); nullptr if that QP is not in the build
const QpLaunchers* qp_launchers(int qp);
#define DPGP_QP_LIST(X) X(2) X(4) X(6) X(8) X(10) X(12) X(16) X(20) X(24) X(28) X(32)
#define DPGP_DECLARE_QP(q) const QpLaunchers* qp_launchers_##q();
DPGP_QP_LIST(DPGP_DECLARE_QP)
#undef DPGP_DECLARE_QP

}  // namespace dpgp
