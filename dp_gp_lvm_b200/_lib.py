"""ctypes binding of the C ABI in include/dpgp.h (libdpgp.so, built in-tree by `make`).

There is no CPU fallback: importing this module without the shared library, or calling into it without a
CUDA device, raises.  The oracle under oracle/ is test infrastructure and is never imported from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DPGP_LIB", os.path.join(_HERE, "libdpgp.so"))      # DPGP_LIB: development override

MODE_T, MODE_D = 0, 1
OK, E_ARG, E_CUDA, E_NOT_PD, E_NOMEM = 0, -1, -2, -3, -4


class DpgpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dpgp error %d: %s" % (code, msg))
        self.code = code


class NotPositiveDefiniteError(DpgpError):
    """Cholesky of K_uu + 1e-8 I or beta H + I failed (the reference's tf.cholesky raises here too)."""


class Options(C.Structure):
    _fields_ = [("exp_variant", C.c_int), ("psi2_threads", C.c_int), ("psi2_chunk", C.c_int),
                ("max_ctas", C.c_int), ("bwd_variant", C.c_int), ("chain_variant", C.c_int), ("reserved", C.c_int * 10)]


class SmallArgs(C.Structure):
    """dpgp_small_args (include/dpgp.h): device pointers of the N-independent variables, outputs and cotangents."""
    _PTRS = ("logits", "gamma1_raw", "gamma2_raw", "w1_raw", "w2_raw", "gamma_atoms_raw", "alpha_atoms_raw", "beta_atoms_raw",
             "phi", "gamma", "alpha", "beta", "scal", "dphi", "dgamma", "dalpha", "dbeta", "grad_out",
             "dlogits", "dgamma1_raw", "dgamma2_raw", "dw1_raw", "dw2_raw", "dgamma_atoms_raw", "dalpha_atoms_raw", "dbeta_atoms_raw")
    _fields_ = [(k, C.c_void_p) for k in _PTRS] + [("truncation_level", C.c_int), ("mask_size", C.c_int),
                                                    ("alpha_prior_shape", C.c_double), ("alpha_prior_rate", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libdpgp.so not found at %s -- build it with `make` (python -c 'import __graft_entry__ as g; "
                          "g.build()'); there is no CPU fallback" % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    dp, vp, i64, ci = C.c_void_p, C.c_void_p, C.c_int64, C.c_int
    l.dpgp_create.argtypes = [C.POINTER(C.c_void_p), ci, i64, ci, ci, ci, ci, ci, C.POINTER(Options)]
    l.dpgp_create.restype = ci
    l.dpgp_destroy.argtypes = [vp]; l.dpgp_destroy.restype = ci
    l.dpgp_last_error.argtypes = [vp]; l.dpgp_last_error.restype = C.c_char_p
    l.dpgp_check.argtypes = [vp, vp]; l.dpgp_check.restype = ci
    l.dpgp_stats_len.argtypes = [vp]; l.dpgp_stats_len.restype = C.c_size_t
    l.dpgp_workspace_bytes.argtypes = [vp]; l.dpgp_workspace_bytes.restype = C.c_size_t
    l.dpgp_launch_count.argtypes = [vp]; l.dpgp_launch_count.restype = i64
    l.dpgp_covariance.argtypes = [vp, dp, i64, dp, i64, dp, dp, dp, ci, ci, dp, vp]; l.dpgp_covariance.restype = ci
    l.dpgp_psi1.argtypes = [vp, dp, dp, i64, dp, dp, dp, dp, vp]; l.dpgp_psi1.restype = ci
    l.dpgp_stats_fwd.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, vp]; l.dpgp_stats_fwd.restype = ci
    l.dpgp_bound.argtypes = [vp, i64] + [dp] * 13 + [vp]; l.dpgp_bound.restype = ci
    l.dpgp_stats_bwd.argtypes = [vp] + [dp] * 12 + [vp]; l.dpgp_stats_bwd.restype = ci
    l.dpgp_set_timing.argtypes = [vp, ci]; l.dpgp_set_timing.restype = ci
    l.dpgp_get_timings.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), ci]; l.dpgp_get_timings.restype = ci
    l.dpgp_bound_factors.argtypes = [vp, dp, dp, dp, vp]; l.dpgp_bound_factors.restype = ci
    l.dpgp_adam.argtypes = [vp, dp, dp, dp, dp, i64, dp, C.c_double, C.c_double, C.c_double, C.c_double, vp]; l.dpgp_adam.restype = ci
    l.dpgp_adam_multi.argtypes = [vp, ci, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_int64), dp, C.c_double, C.c_double, C.c_double, C.c_double, vp]; l.dpgp_adam_multi.restype = ci
    l.dpgp_train_tail.argtypes = [vp, dp, dp, dp, ci, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_void_p), dp, C.c_double, C.c_double, C.c_double,
                                  C.c_double, vp]; l.dpgp_train_tail.restype = ci
    l.dpgp_fused_schedule.argtypes = [ci, C.POINTER(C.c_ushort), ci]; l.dpgp_fused_schedule.restype = ci
    l.dpgp_small_fwd.argtypes = [vp, C.POINTER(SmallArgs), vp]; l.dpgp_small_fwd.restype = ci
    l.dpgp_small_bwd.argtypes = [vp, C.POINTER(SmallArgs), vp]; l.dpgp_small_bwd.restype = ci
    l.dpgp_polygamma.argtypes = [dp, dp, dp, i64, vp]; l.dpgp_polygamma.restype = ci
    l.dpgp_debug_launch_times.argtypes = [vp, ci, C.POINTER(C.c_char_p), C.POINTER(C.c_float), ci]; l.dpgp_debug_launch_times.restype = ci
    l.dpgp_check_guards.argtypes = [vp]; l.dpgp_check_guards.restype = ci
    l.dpgp_has_experimental.argtypes = []; l.dpgp_has_experimental.restype = ci
    l.dpgp_limits.argtypes = [C.POINTER(ci), C.POINTER(ci)]; l.dpgp_limits.restype = ci
    _lib = l
    return l


EXPORTS = ("dpgp_create", "dpgp_destroy", "dpgp_last_error", "dpgp_check", "dpgp_stats_len", "dpgp_workspace_bytes",
           "dpgp_launch_count", "dpgp_covariance", "dpgp_psi1", "dpgp_stats_fwd", "dpgp_bound", "dpgp_stats_bwd",
           "dpgp_set_timing", "dpgp_get_timings", "dpgp_fused_schedule", "dpgp_adam", "dpgp_bound_factors",
           "dpgp_small_fwd", "dpgp_small_bwd", "dpgp_adam_multi", "dpgp_train_tail", "dpgp_has_experimental", "dpgp_limits", "dpgp_polygamma", "dpgp_debug_launch_times", "dpgp_check_guards")


def has_experimental():
    """True if libdpgp.so was built with `make EXPERIMENTAL=1` (csrc/experimental/ variants selectable)."""
    return bool(lib().dpgp_has_experimental())


def limits():
    """(max Q, max M) of this build."""
    q, m = C.c_int(), C.c_int()
    lib().dpgp_limits(C.byref(q), C.byref(m))
    return q.value, m.value
