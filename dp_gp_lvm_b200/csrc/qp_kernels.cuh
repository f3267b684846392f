// Per-latent-dimension kernel launchers.  qp_kernels.cu is compiled once per padded latent dimension
// (-DDPGP_QP=2,4,...,16) so the seven instantiation sets build in parallel; dpgp_api.cu selects the
// table for the handle's QP.
#pragma once
#include <cuda_runtime.h>

#include "chain.cuh"
#include "chain2.cuh"
#include "psi1.cuh"
#include "psi2.cuh"
#include "psi2_bwd.cuh"
#include "psi2_bwd_fused.cuh"
#include "psi2_bwd_tc.cuh"
#include "psi2_bwd_ws.cuh"

namespace dpgp {

struct QpLaunchers {
  cudaError_t (*cfg_smem)(int expv, size_t f, size_t pp, size_t nn, size_t p1, size_t g1, size_t ch, int urows, size_t fused);
  size_t (*fused_smem)(int rows, int mp);
  void (*psi2_bwd_fused)(int expv, int rows, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool dz);
  // two teams of 8 warps per CTA on 32-row groups (16 warps / SM); configure_only sets the shared-memory attribute
  size_t (*fused2_smem)(int mp);
  bool (*psi2_bwd_fused2)(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only);
  // warp-specialised: 8 producer warps (phase 1) + 8 helper warps (phase 2), setmaxnreg
  size_t (*ws_smem)(int mp);
  bool (*psi2_bwd_ws)(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only);
  // tensor-core formulation of the first phase (QP <= 12, 64-row groups); returns false if not instantiated for this QP
  bool (*psi2_bwd_tc)(int expv, int grid, size_t smem, cudaStream_t st, const Psi2BwdFusedParams& p, bool configure_only);
  void (*prep)(int grid, cudaStream_t st, const PrepParams& p);
  void (*psi2_fwd)(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2FwdParams& p);
  void (*psi2_bwd_pair)(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2BwdPairParams& p);
  void (*psi2_bwd_n)(int expv, int grid, int threads, size_t smem, cudaStream_t st, const Psi2BwdNParams& p);
  void (*psi1_fwd)(int grid, size_t smem, cudaStream_t st, const Psi1FwdParams& p);
  void (*g1)(int grid, size_t smem, cudaStream_t st, const G1Params& p);
  void (*chain)(int grid, size_t smem, cudaStream_t st, const ChainParams& p);
  size_t (*chain2_smem)(int rows, int mp);
  cudaError_t (*chain2_cfg)(int rows, size_t smem);
  void (*chain2)(int rows, int grid, size_t smem, cudaStream_t st, const Chain2Params& p);
};

const QpLaunchers* qp_launchers_2();
const QpLaunchers* qp_launchers_4();
const QpLaunchers* qp_launchers_6();
const QpLaunchers* qp_launchers_8();
const QpLaunchers* qp_launchers_10();
const QpLaunchers* qp_launchers_12();
const QpLaunchers* qp_launchers_16();

}  // namespace dpgp
