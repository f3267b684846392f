/* dpgp.h -- C ABI of the B200-native DP-GP-LVM bound (ELBO + gradients) hot path.
 *
 * The reference (AndrewRLawrence/dp_gp_lvm) has no FFI boundary: its hot path is a TensorFlow-1 graph
 * built by  src/kernels/rbf_kernel.py:26-203 (k_ard_rbf: K_uu, psi_0/1/2) and
 * src/models/dp_gp_lvm.py:100-154 (D-mode bound) / :582-676 (T-mode bound), differentiated by
 * tf.gradients (test/synthetic_data_hard_test.py:143).  This header is the boundary a maintainer
 * binds instead (ctypes stub in INTEGRATION.md); every entry point states the reference lines it
 * replaces.
 *
 * Conventions
 *  - float64 everywhere (src/utils/types.py:13-14), row-major, contiguous.
 *  - Every pointer argument named d_* is a DEVICE pointer owned by the caller; the library neither
 *    frees nor retains it beyond the call.  Scratch lives in the handle.
 *  - All calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *    and return 0 on success or a negative DPGP_E_* code; no C++ exception crosses this boundary.
 *    Numerical failures detected on the device (non-positive Cholesky pivot) are reported by
 *    dpgp_check(), which synchronises the stream.
 *  - A handle is bound to one device and one (N_local, D, Q, M, B, mode) shape; it is re-entrant
 *    across handles and not thread-safe on one handle.
 *
 * Kernel batch B: T-mode (dp_gp_lvm_t) B = T clusters and every kernel sees all D columns of Y,
 * weighted by phi[d,b]; D-mode (dp_gp_lvm) B = D and kernel b sees only column b with weight 1.
 *
 * Packed statistics buffer (one NCCL all-reduce covers it), dpgp_stats_len() doubles:
 *    [ Psi2 : B*M*M | P : B*M*C | yy : D | kl : 2 ]      C = D (T-mode) or 1 (D-mode)
 *    Psi2[b] = sum_n psi2_n            (rbf_kernel.py:189-199)
 *    P[b]    = Psi1[b]^T Y(:,cols(b))  (rbf_kernel.py:155-161 contracted as dp_gp_lvm.py:638-658 needs)
 *    yy[d]   = sum_n y_nd^2            (dp_gp_lvm.py:143 / :659)
 *    kl      = { sum mu^2 , sum (s - log s) }   (gp_expressions.py:18-23)
 */
#ifndef DPGP_H_
#define DPGP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dpgp_handle dpgp_handle;

enum { DPGP_MODE_T = 0, DPGP_MODE_D = 1 };

enum {
  DPGP_OK = 0,
  DPGP_E_ARG = -1,        /* bad shape / null pointer / unsupported size */
  DPGP_E_CUDA = -2,       /* a CUDA runtime call failed (see dpgp_last_error) */
  DPGP_E_NOT_PD = -3,     /* Cholesky met a non-positive pivot (K_uu + 1e-8 I or beta H + I) */
  DPGP_E_NOMEM = -4
};

/* Tunables, all optional (0 = library default).  The default build of the library holds the default kernels and one
 * fallback each (exp_variant 0 = 4 and 1; bwd_variant 0 = 6, plus 1, 7 and 8; chain_variant 0 = 1); the other values name the
 * experimental variants under csrc/experimental/, built only by `make EXPERIMENTAL=1` (dpgp_has_experimental()), and are
 * rejected with DPGP_E_ARG otherwise. */
typedef struct dpgp_options {
  int exp_variant;        /* 0 default (shared-memory table), 1 libdevice exp, 2 poly11, 3 shuffle-table,
                             4 / 5 / 6: 256- / 64- / 32-entry shared-memory table + degree-4 / 5 / 6 polynomial */
  int psi2_threads;       /* CTA size of the psi2 forward kernel (multiple of 32) */
  int psi2_chunk;         /* rows of q(X) staged per shared-memory tile */
  int max_ctas;           /* persistent grid size (default: number of SMs) */
  int bwd_variant;        /* psi2 backward: 0 default (= 6), 1 fused (one exp per unit), dD in per-CTA slices, 2 two-kernel (pair + row),
                             3 fused with the first phase on the FP64 tensor cores (QP <= 12),
                             4 fused with two 8-warp teams per CTA on 32-row groups (16 warps / SM),
                             5 fused, warp-specialised: 8 producer warps (phase 1) + 8 helper warps (phase 2) per CTA,
                               mbarrier hand-over of the g tiles, setmaxnreg 200 / 56 (QP <= 12),
                             6 fused with the pair-side totals folded into dZ inside the kernel: 24 MB of per-warp slices
                               instead of ~200 MB of per-CTA dD slices, 6x less DRAM traffic, 1 % slower than 1,
                             7 as 6 with dv and dD on the tcgen05 tensor cores: g written once as six int8 planes, slice
                               products with int32 accumulators in TMEM (csrc/psi2_bwd_umma.cuh; Q <= 16, M <= 128; exact to 6e-16, 1.2x slower),
                             8 as 6 with the dv / dD contractions as FP64 DMMA on q < 8 (csrc/psi2_bwd_mma.cuh; 7 <= Q <= 12; same speed) */
  int chain_variant;      /* psi1 backward + chain: 0 default (fused, FP64 tensor-core contractions), 1 same, 2 two-kernel */
  int reserved[10];
} dpgp_options;

/* 1 if the library was built with `make EXPERIMENTAL=1` (non-default kernel variants available), else 0.  Host-only. */
int dpgp_has_experimental(void);
/* Compile-time size limits of this build (host-only): latent dimension Q and inducing points M. */
int dpgp_limits(int* max_q, int* max_m);

/* Creates a handle: allocates workspace for n_local rows on `device`.  mode = DPGP_MODE_T/D.
 * Limits: 1 <= Q <= 32 (the reference's scripts use up to 25), 1 <= M <= 256, B >= 1, D >= 1; the reference itself
 * (src/kernels/rbf_kernel.py:26-45) is unbounded.  Large (Q, M) combinations are additionally bounded by the 227 KB of
 * shared memory per CTA (e.g. M <= 256 up to Q = 12, M <= 128 at Q = 32): dpgp_create then fails with DPGP_E_ARG and a message naming the kernel. */
int dpgp_create(dpgp_handle** out, int device, int64_t n_local, int d, int q, int m, int b, int mode,
                const dpgp_options* opt /* may be NULL */);
int dpgp_destroy(dpgp_handle* h);

/* Message of the most recent failing call on this handle ("" if none). */
const char* dpgp_last_error(const dpgp_handle* h);
/* Synchronises `stream` and returns DPGP_E_NOT_PD (with pivot location in dpgp_last_error) if a
 * device-side numerical failure was flagged since the last check, else DPGP_OK. */
int dpgp_check(dpgp_handle* h, void* stream);

size_t dpgp_stats_len(const dpgp_handle* h);
size_t dpgp_workspace_bytes(const dpgp_handle* h);
/* Number of kernel launches issued through this handle since creation (bench.py "gpu_launches"). */
int64_t dpgp_launch_count(const dpgp_handle* h);

/* --- kernel-level entry points: the reference's Kernel object (src/kernels/interfaces/kernel.py:180-280)
 *     evaluated for a batch of B kernels; outputs are caller-owned device buffers. ------------------- */

/* K[b,i,j] = alpha_b exp(-1/2 sum_q gamma_bq (x0_iq - x1_jq)^2); d_x1 == NULL means x1 = x0 and then
 * noise (1/beta_b) / jitter (1e-8) are added on the diagonal if requested (rbf_kernel.py:58-93). */
int dpgp_covariance(dpgp_handle* h, const double* d_x0, int64_t n0, const double* d_x1, int64_t n1,
                    const double* d_gamma, const double* d_alpha, const double* d_beta,
                    int include_noise, int include_jitter, double* d_out /* [B,n0,n1] */, void* stream);

/* Psi1[b,n,m] materialised (rbf_kernel.py:135-161); for the API surface and tests, not used by the bound. */
int dpgp_psi1(dpgp_handle* h, const double* d_mu, const double* d_s, int64_t n, const double* d_z,
              const double* d_gamma, const double* d_alpha, double* d_out /* [B,n,M] */, void* stream);

/* --- the hot path ---------------------------------------------------------------------------------- */

/* Local sufficient statistics of this rank's rows (psi_1, psi_2, Y^T Y diagonal, KL sums), packed as
 * described above.  d_mu, d_s: [N_local,Q]; d_y: [N_local,D]; d_z: [M,Q]; d_gamma: [B,Q]; d_alpha: [B].
 * Replaces rbf_kernel.py:135-199 + the N-contractions of dp_gp_lvm.py:132-145 / :638-667. */
int dpgp_stats_fwd(dpgp_handle* h, const double* d_mu, const double* d_s, const double* d_y,
                   const double* d_z, const double* d_gamma, const double* d_alpha,
                   double* d_stats, void* stream);

/* The M x M chain on the (all-reduced) statistics, forward and backward in one call:
 *   L = chol(K_uu + 1e-8 I); H = L^-1 Psi2 L^-T; A = beta H + I; L_A = chol(A); C = L_A^-1 L^-1 P
 *   (dp_gp_lvm.py:113-145 / :618-667), value  *d_gp = f_hat - KL,
 * and the cotangents of  +(f_hat - KL)  (the caller negates: the objective is  dp - (f_hat - KL) - prior)  w.r.t.
 *   the packed statistics (d_dstats, same layout), K_uu-path and direct parts of Z/gamma/alpha
 *   (d_dz [M,Q], d_dgamma [B,Q], d_dalpha [B]), beta (d_dbeta [B]) and the weights (d_dwgt [D,B] in
 *   T-mode = d f_hat / d phi; ignored (may be NULL) in D-mode).
 * n_total: global number of rows (sum over ranks).  d_wgt: phi [D,B] (T-mode) or NULL (D-mode).
 * Runs replicated and bit-identical on every rank. */
int dpgp_bound(dpgp_handle* h, int64_t n_total, const double* d_stats, const double* d_z,
               const double* d_gamma, const double* d_alpha, const double* d_beta, const double* d_wgt,
               double* d_gp /* [1] */, double* d_dstats, double* d_dz, double* d_dgamma, double* d_dalpha,
               double* d_dbeta, double* d_dwgt, void* stream);

/* Factors of the most recent dpgp_bound call on this handle, for the prediction paths (reference
 * src/models/dp_gp_lvm.py:338-345, :417-500 builds them from L_uu^-1 and L_A^-1): per kernel-batch entry
 *   d_kinv [B,M,M] = (K_uu + 1e-8 I)^-1,  d_sinv [B,M,M] = (K_uu + 1e-8 I + beta Psi2)^-1,  d_u [B,M,C] = sinv P.
 * Any of the three may be NULL.  Stream-ordered after that dpgp_bound. */
int dpgp_bound_factors(dpgp_handle* h, double* d_kinv, double* d_sinv, double* d_u, void* stream);

/* Backward of dpgp_stats_fwd for this rank's rows, given the cotangents d_dstats of (f_hat - KL):
 *   d_dmu, d_ds [N_local,Q]  (complete, incl. the KL term; stay sharded)
 *   d_dz [M,Q], d_dgamma [B,Q], d_dalpha [B]  (this rank's partial sums: all-reduce, then add to the
 *   dpgp_bound outputs).  Replaces TensorFlow autodiff through rbf_kernel.py:135-199. */
int dpgp_stats_bwd(dpgp_handle* h, const double* d_mu, const double* d_s, const double* d_y,
                   const double* d_z, const double* d_gamma, const double* d_alpha,
                   const double* d_dstats, double* d_dmu, double* d_ds, double* d_dz, double* d_dgamma,
                   double* d_dalpha, void* stream);

/* Per-phase device timings (ms) of the most recent stats_fwd / bound / stats_bwd calls, measured with
 * CUDA events on `stream` when timing is enabled.  names: "prep","psi2_fwd","psi1_fwd","bound",
 * "psi2_bwd_n","psi2_bwd_pair","chain_bwd","reduce","psi2_bwd_fused".  Returns the number of entries written (<= cap). */
int dpgp_set_timing(dpgp_handle* h, int enabled);
int dpgp_get_timings(dpgp_handle* h, const char** names, float* ms, int cap);

/* Development aid: if the environment variable DPGP_GUARD was set when the handle was created, every workspace buffer of the
 * handle sits between two 4 KB bands of a fixed byte pattern.  Synchronises the device and returns the number of buffers
 * whose bands were overwritten (0 = no kernel wrote out of bounds next to the workspace; dpgp_last_error names the first
 * offender), or DPGP_E_ARG without DPGP_GUARD. */
int dpgp_check_guards(dpgp_handle* h);

/* Development aid: with enable != 0, every kernel launched through this handle from now on is followed by an event on the
 * stream of the surrounding hot-path call; the next call of this function synchronises the device, writes the name and the
 * device time (microseconds, previous event -> this event, i.e. including any wait on the stream) of each launch recorded
 * since, returns their number (<= cap) and clears the record.  Side-stream launches (the K_uu factor) are timed on the main
 * stream's clock and therefore show the gap they leave, not their own duration.  Not for use under CUDA-graph capture. */
int dpgp_debug_launch_times(dpgp_handle* h, int enable, const char** names, float* us, int cap);

/* --- the caller of the hot path: one optimiser step ------------------------------------
 * Adam update of one flat parameter tensor in TensorFlow-1's formulation, as tf.train.AdamOptimizer(lr)
 * .minimize(objective) applies it in every reference script (test/synthetic_data_hard_test.py:143,152):
 *   lr_t = lr sqrt(1 - beta2^t) / (1 - beta1^t);  m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2;
 *   param -= lr_t m / (sqrt(v) + eps)
 * d_step: device int64 holding t >= 1 (incremented by the caller once per iteration, on the same stream), so that
 * a CUDA-graph capture of the whole training iteration stays valid.  d_m / d_v: the optimiser slots (zero at t = 0). */
int dpgp_adam(dpgp_handle* h, double* d_param, const double* d_grad, double* d_m, double* d_v, int64_t n,
              const int64_t* d_step, double lr, double beta1, double beta2, double eps, void* stream);
/* The same update for `count` tensors in one launch (host arrays of device pointers and element counts). */
int dpgp_adam_multi(dpgp_handle* h, int count, double* const* d_params, const double* const* d_grads, double* const* d_ms,
                    double* const* d_vs, const int64_t* ns, const int64_t* d_step, double lr, double beta1, double beta2,
                    double eps, void* stream);

/* --- the N-independent part of the objective, fused ("softplus / softmax chain" on device) --------
 * Replaces the few hundred TensorFlow ops per iteration of src/models/dirichlet_process.py:33-88 (phi = softmax(logits)
 * with mask_size tying, q(V) / q(alpha) parameters, the six ELBO terms), the softplus-positive atoms
 * (src/utils/types.py:52-57), their log-normal hyper-prior (dp_gp_lvm.py:96-98 / :603-605) and, in D-mode, the
 * phi-mixtures of the atoms (dp_gp_lvm.py:100-102) -- and their autodiff -- by one forward and one backward launch.
 * All pointers are device pointers to contiguous float64; the handle's D, Q, B and mode apply.
 *   forward : raw variables -> phi [D,T], gamma [B,Q], alpha [B], beta [B], scal[0] = DP objective (-ELBO), scal[1] = prior
 *   backward: cotangents of the GP bound (dphi [D,T] in T-mode / NULL in D-mode, dgamma [B,Q], dalpha [B], dbeta [B]) ->
 *             gradients of  objective = scal[0] - gp - scal[1]  w.r.t. every raw variable, times *grad_out (NULL = 1). */
typedef struct dpgp_small_args {
  const double* logits;            /* [D / mask_size, T] */
  const double* gamma1_raw; const double* gamma2_raw;     /* [T-1] */
  const double* w1_raw; const double* w2_raw;             /* scalars */
  const double* gamma_atoms_raw;   /* [T,Q] */
  const double* alpha_atoms_raw; const double* beta_atoms_raw;   /* [T] */
  double* phi; double* gamma; double* alpha; double* beta; double* scal;            /* forward outputs (backward: phi is input) */
  const double* dphi; const double* dgamma; const double* dalpha; const double* dbeta; const double* grad_out;
  double* dlogits; double* dgamma1_raw; double* dgamma2_raw; double* dw1_raw; double* dw2_raw;
  double* dgamma_atoms_raw; double* dalpha_atoms_raw; double* dbeta_atoms_raw;
  int truncation_level; int mask_size;
  double alpha_prior_shape; double alpha_prior_rate;      /* (s_1, s_2) of dirichlet_process(alpha_prior_params) */
} dpgp_small_args;
int dpgp_small_fwd(dpgp_handle* h, const dpgp_small_args* a, void* stream);
int dpgp_small_bwd(dpgp_handle* h, const dpgp_small_args* a, void* stream);

/* The tail of one optimiser iteration without any host-framework glue (two launches; the reference's
 * `session.run(AdamOptimizer(...).minimize(model.objective))`, test/synthetic_data_hard_test.py:143-155):
 *   *d_objective = d_scal[0] - d_scal[1] - *d_gp      (dp_gp_lvm.py:148-154 / :670-676; d_scal from dpgp_small_fwd, d_gp from dpgp_bound)
 *   ++*d_step
 *   dpgp_adam_multi with the gradient of tensor i taken as  g_scale[i] * d_grads[i] * (d_raws[i] ? sigmoid(d_raws[i]) : 1):
 *   g_scale (host array) carries the sign with which a bound gradient enters the objective (-1 for q(X) and the inducing inputs),
 *   d_raws[i] != NULL applies the chain rule of a softplus-parameterised variable (x_var = softplus(raw), src/utils/types.py:52-57)
 *   to a gradient that was taken with respect to the positive value.  d_raws may be NULL. */
int dpgp_train_tail(dpgp_handle* h, const double* d_scal, const double* d_gp, double* d_objective, int count, double* const* d_params,
                    const double* const* d_grads, double* const* d_ms, double* const* d_vs, const int64_t* ns, const double* g_scale,
                    const double* const* d_raws, int64_t* d_step, double lr, double beta1, double beta2, double eps, void* stream);

/* Elementwise digamma psi(x) and trigamma psi'(x) for x > 0 (NaN otherwise) with the device functions the fused kernels
 * use (csrc/special.cuh, ~1e-16 relative): the stand-alone `dirichlet_process` model (src/models/dirichlet_process.py:64-77,
 * src/distributions/beta.py:18-19, gamma.py:17) differentiates tf.digamma; torch's own trigamma is only good to ~5e-10 in
 * float64, which the cancellation in d ELBO / d w_1 amplifies beyond the 1e-9 parity bar.  Either output may be NULL.
 * No handle: launches on the current device. */
int dpgp_polygamma(const double* d_x, double* d_digamma, double* d_trigamma, int64_t n, void* stream);

/* Host-only helper (no GPU needed): the block schedule of the fused psi2 backward kernel for `num_mblocks`
 * = ceil(M/8) blocks of 8 inducing points.  Writes rounds x 8 entries ((bi << 8) | bj, 0xffff = idle warp)
 * into out[0..cap) and returns the number of rounds (< 0 on bad arguments).  Within a round no two entries
 * share an m-block, and every block bi <= bj appears exactly once overall: this is what makes the shared
 * d r accumulation of that kernel conflict-free and deterministic (tests/test_abi.py checks it). */
int dpgp_fused_schedule(int num_mblocks, unsigned short* out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* DPGP_H_ */
