"""Thin, torch-tensor-facing wrapper of one `dpgp_handle` (include/dpgp.h).

PyTorch is used only for device memory, streams and (in models/) torch.distributed; every number comes
from the CUDA kernels behind the C ABI.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import MODE_D, MODE_T, DpgpError, NotPositiveDefiniteError, Options


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous(), "need contiguous float64 CUDA tensors"
    return C.c_void_p(t.data_ptr())


# dpgp_destroy synchronises the handle's side stream and frees its workspace: both are illegal while a CUDA-graph capture is
# under way (they invalidate the capture).  An engine can die at any moment -- Python's cycle collector runs whenever it likes,
# e.g. in the middle of somebody's torch.cuda.graph block -- so handles released during a capture are parked here and destroyed
# by the next create / close outside a capture.
_DEFERRED_DESTROY = []


def _capturing():
    try:
        return torch.cuda.is_current_stream_capturing()
    except Exception:
        return False


def _drain_deferred(lib):
    while _DEFERRED_DESTROY and not _capturing():
        lib.dpgp_destroy(_DEFERRED_DESTROY.pop())


class BoundEngine:
    """Workspace + kernels for one (N_local, D, Q, M, B, mode) shape on one GPU."""

    def __init__(self, n_local, d, q, m, b, mode, device=None, exp_variant=0, psi2_threads=0, psi2_chunk=0, max_ctas=0,
                 bwd_variant=0, chain_variant=0):
        if not torch.cuda.is_available():
            raise RuntimeError("dp_gp_lvm_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        _drain_deferred(self.lib)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:                       # "cuda" means the CURRENT device, not device 0
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n, self.d, self.q, self.m, self.b, self.mode = int(n_local), int(d), int(q), int(m), int(b), int(mode)
        self.ncols = self.d if self.mode == MODE_T else 1
        opt = Options(exp_variant=exp_variant, psi2_threads=psi2_threads, psi2_chunk=psi2_chunk, max_ctas=max_ctas,
                      bwd_variant=bwd_variant, chain_variant=chain_variant)
        self._h = C.c_void_p()
        rc = self.lib.dpgp_create(C.byref(self._h), self.device.index, self.n, self.d, self.q, self.m, self.b,
                                  self.mode, C.byref(opt))
        if rc != 0:
            msg = self.lib.dpgp_last_error(self._h).decode() if self._h else "allocation failed"
            if self._h:
                self.lib.dpgp_destroy(self._h)
            self._h = None
            raise DpgpError(rc, msg)
        self.stats_len = int(self.lib.dpgp_stats_len(self._h))

    # -- plumbing ---------------------------------------------------------------------------------
    def close(self):
        h = getattr(self, "_h", None)
        if h:
            self._h = None
            if _capturing():
                _DEFERRED_DESTROY.append(h)
            else:
                self.lib.dpgp_destroy(h)
                _drain_deferred(self.lib)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc):
        if rc != 0:
            msg = self.lib.dpgp_last_error(self._h).decode()
            raise (NotPositiveDefiniteError if rc == _lib.E_NOT_PD else DpgpError)(rc, msg)

    def check(self):
        """Synchronises and raises NotPositiveDefiniteError if a Cholesky pivot was non-positive."""
        self._ck(self.lib.dpgp_check(self._h, self._stream()))

    @property
    def workspace_bytes(self):
        return int(self.lib.dpgp_workspace_bytes(self._h))

    @property
    def launch_count(self):
        return int(self.lib.dpgp_launch_count(self._h))

    def set_timing(self, enabled):
        self.lib.dpgp_set_timing(self._h, int(bool(enabled)))

    def timings(self):
        names = (C.c_char_p * 16)(); ms = (C.c_float * 16)()
        k = self.lib.dpgp_get_timings(self._h, names, ms, 16)
        return {names[i].decode(): float(ms[i]) for i in range(k)}

    def check_guards(self):
        """Development aid (dpgp_check_guards; needs DPGP_GUARD in the environment when the engine was created): raises if a
        kernel wrote outside a workspace buffer."""
        rc = self.lib.dpgp_check_guards(self._h)
        if rc != 0:
            raise DpgpError(rc, self.lib.dpgp_last_error(self._h).decode())

    def launch_times(self, enable=True):
        """Development aid (dpgp_debug_launch_times): [(kernel name, microseconds)] of the launches recorded since the previous
        call; `enable` switches the recording on or off for the calls that follow."""
        names = (C.c_char_p * 512)(); us = (C.c_float * 512)()
        k = self.lib.dpgp_debug_launch_times(self._h, int(bool(enable)), names, us, 512)
        return [(names[i].decode(), float(us[i])) for i in range(k)]

    def _new(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    # -- views into the packed statistics buffer ----------------------------------------------------
    def split_stats(self, stats):
        b, m, c, d = self.b, self.m, self.ncols, self.d
        o1 = b * m * m; o2 = o1 + b * m * c; o3 = o2 + d
        return (stats[:o1].view(b, m, m), stats[o1:o2].view(b, m, c), stats[o2:o3], stats[o3:o3 + 2])

    # -- kernel-level API ------------------------------------------------------------------------------
    def covariance(self, x0, x1, gamma, alpha, beta, include_noise=False, include_jitter=False):
        n0 = x0.shape[0]; n1 = n0 if x1 is None else x1.shape[0]
        out = self._new(self.b, n0, n1)
        self._ck(self.lib.dpgp_covariance(self._h, _ptr(x0), n0, _ptr(x1), n1, _ptr(gamma), _ptr(alpha), _ptr(beta),
                                          int(include_noise), int(include_jitter), _ptr(out), self._stream()))
        return out

    def psi1(self, mu, s, z, gamma, alpha):
        n = mu.shape[0]
        assert n == self.n, "handle was created for a different number of rows"
        out = self._new(self.b, n, self.m)
        self._ck(self.lib.dpgp_psi1(self._h, _ptr(mu), _ptr(s), n, _ptr(z), _ptr(gamma), _ptr(alpha), _ptr(out), self._stream()))
        return out

    # -- hot path -----------------------------------------------------------------------------------------
    def stats_fwd(self, mu, s, y, z, gamma, alpha, out=None):
        assert mu.shape == (self.n, self.q) and s.shape == (self.n, self.q) and y.shape == (self.n, self.d)
        assert z.shape == (self.m, self.q) and gamma.shape == (self.b, self.q) and alpha.numel() == self.b
        stats = self._new(self.stats_len) if out is None else out
        self._ck(self.lib.dpgp_stats_fwd(self._h, _ptr(mu), _ptr(s), _ptr(y), _ptr(z), _ptr(gamma), _ptr(alpha), _ptr(stats), self._stream()))
        return stats

    def small_views(self, buf):
        """(dz [M,Q], dgamma [B,Q], dalpha [B]) as views of one contiguous buffer of M Q + B Q + B doubles."""
        nz, ng = self.m * self.q, self.b * self.q
        return buf[:nz].view(self.m, self.q), buf[nz:nz + ng].view(self.b, self.q), buf[nz + ng:nz + ng + self.b]

    def bound(self, n_total, stats, z, gamma, alpha, beta, wgt, small_out=None):
        """Returns (gp [1], dstats, dz, dgamma, dalpha, dbeta, dwgt) -- cotangents of gp = f_hat - KL.  `small_out`: a buffer of
        M Q + B Q + B doubles that receives [dz | dgamma | dalpha] contiguously (the returned tensors are then views of it)."""
        gp = self._new(1); dstats = self._new(self.stats_len); dbeta = self._new(self.b)
        if small_out is None:
            dz = self._new(self.m, self.q); dgamma = self._new(self.b, self.q); dalpha = self._new(self.b)
        else:
            dz, dgamma, dalpha = self.small_views(small_out)
        dwgt = self._new(self.d, self.b) if self.mode == MODE_T else None
        self._ck(self.lib.dpgp_bound(self._h, int(n_total), _ptr(stats), _ptr(z), _ptr(gamma), _ptr(alpha), _ptr(beta),
                                     _ptr(wgt) if self.mode == MODE_T else None, _ptr(gp), _ptr(dstats), _ptr(dz),
                                     _ptr(dgamma), _ptr(dalpha), _ptr(dbeta), _ptr(dwgt), self._stream()))
        return gp, dstats, dz, dgamma, dalpha, dbeta, dwgt

    def bound_factors(self):
        """(K^-1 [B,M,M], (K + beta Psi2)^-1 [B,M,M], (K + beta Psi2)^-1 P [B,M,C]) of the most recent bound() call."""
        kinv = self._new(self.b, self.m, self.m); sinv = self._new(self.b, self.m, self.m); u = self._new(self.b, self.m, self.ncols)
        self._ck(self.lib.dpgp_bound_factors(self._h, _ptr(kinv), _ptr(sinv), _ptr(u), self._stream()))
        return kinv, sinv, u

    def stats_bwd(self, mu, s, y, z, gamma, alpha, dstats, small_out=None):
        dmu = self._new(self.n, self.q); ds = self._new(self.n, self.q)
        if small_out is None:
            dz = self._new(self.m, self.q); dgamma = self._new(self.b, self.q); dalpha = self._new(self.b)
        else:
            dz, dgamma, dalpha = self.small_views(small_out)
        self._ck(self.lib.dpgp_stats_bwd(self._h, _ptr(mu), _ptr(s), _ptr(y), _ptr(z), _ptr(gamma), _ptr(alpha), _ptr(dstats),
                                         _ptr(dmu), _ptr(ds), _ptr(dz), _ptr(dgamma), _ptr(dalpha), self._stream()))
        return dmu, ds, dz, dgamma, dalpha

    # -- the N-independent part of the objective (include/dpgp.h: dpgp_small_fwd / dpgp_small_bwd) ----------------
    def _small_args(self, raw, truncation_level, mask_size, alpha_prior, **ptrs):
        a = _lib.SmallArgs()
        for k, t in raw.items():
            setattr(a, k, None if t is None else _ptr(t))
        for k, t in ptrs.items():
            setattr(a, k, None if t is None else _ptr(t))
        a.truncation_level = int(truncation_level); a.mask_size = int(mask_size)
        a.alpha_prior_shape = float(alpha_prior[0]); a.alpha_prior_rate = float(alpha_prior[1])
        return a

    def small_fwd(self, raw, truncation_level, mask_size, alpha_prior):
        """raw: dict of the raw variables (keys of dpgp_small_args).  Returns (phi [D,T], gamma [B,Q], alpha [B], beta [B],
        scal [2] = (DP objective, hyper-prior))."""
        phi = self._new(self.d, truncation_level); gamma = self._new(self.b, self.q); alpha = self._new(self.b)
        beta = self._new(self.b); scal = self._new(2)
        a = self._small_args(raw, truncation_level, mask_size, alpha_prior, phi=phi, gamma=gamma, alpha=alpha, beta=beta, scal=scal)
        self._ck(self.lib.dpgp_small_fwd(self._h, C.byref(a), self._stream()))
        return phi, gamma, alpha, beta, scal

    def small_bwd(self, raw, truncation_level, mask_size, alpha_prior, phi, dphi, dgamma, dalpha, dbeta, out, grad_out=None):
        """out: dict of gradient tensors keyed dlogits, dgamma1_raw, ... (written)."""
        a = self._small_args(raw, truncation_level, mask_size, alpha_prior, phi=phi, dphi=dphi, dgamma=dgamma, dalpha=dalpha,
                             dbeta=dbeta, grad_out=grad_out, **out)
        self._ck(self.lib.dpgp_small_bwd(self._h, C.byref(a), self._stream()))

    # -- optimiser step (the caller of the hot path) ---------------------------------------------------------
    def adam(self, param, grad, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        """In-place TensorFlow-1 Adam update of `param` (include/dpgp.h: dpgp_adam); `step` is a device int64 tensor."""
        assert param.is_contiguous() and grad.is_contiguous() and m.is_contiguous() and v.is_contiguous()
        assert step.dtype == torch.int64 and step.is_cuda
        self._ck(self.lib.dpgp_adam(self._h, _ptr(param), _ptr(grad), _ptr(m), _ptr(v), param.numel(),
                                    C.c_void_p(step.data_ptr()), float(lr), float(beta1), float(beta2), float(eps), self._stream()))

    def train_tail(self, scal, gp, objective_out, params, grads, ms, vs, scales, raws, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        """dpgp_train_tail: objective = scal[0] - scal[1] - gp, ++step, and the Adam update of every tensor with gradient
        scales[i] * grads[i] (* sigmoid(raws[i]) where raws[i] is not None), in two launches."""
        assert step.dtype == torch.int64 and step.is_cuda and objective_out.is_cuda and objective_out.dtype == torch.float64
        k = len(params)
        arr = lambda ts: (C.c_void_p * k)(*[None if t is None else t.data_ptr() for t in ts])
        for t in list(params) + list(grads) + list(ms) + list(vs) + [r for r in raws if r is not None]:
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        ns = (C.c_int64 * k)(*[p.numel() for p in params])
        sc = (C.c_double * k)(*[float(x) for x in scales])
        self._ck(self.lib.dpgp_train_tail(self._h, _ptr(scal), _ptr(gp), C.c_void_p(objective_out.data_ptr()), k, arr(params), arr(grads), arr(ms),
                                          arr(vs), ns, sc, arr(raws), C.c_void_p(step.data_ptr()), float(lr), float(beta1), float(beta2),
                                          float(eps), self._stream()))

    def adam_multi(self, params, grads, ms, vs, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        """dpgp_adam_multi: the same update for a list of tensors in one launch."""
        assert step.dtype == torch.int64 and step.is_cuda
        k = len(params)
        arr = lambda ts: (C.c_void_p * k)(*[t.data_ptr() for t in ts])
        for t in list(params) + list(grads) + list(ms) + list(vs):
            assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        ns = (C.c_int64 * k)(*[p.numel() for p in params])
        self._ck(self.lib.dpgp_adam_multi(self._h, k, arr(params), arr(grads), arr(ms), arr(vs), ns, C.c_void_p(step.data_ptr()),
                                          float(lr), float(beta1), float(beta2), float(eps), self._stream()))
