"""digamma with an accurate derivative (reference: tf.digamma in src/models/dirichlet_process.py:64-77,
src/distributions/beta.py:18-19, gamma.py:17, differentiated by tf.gradients).

torch.digamma's autograd derivative (torch.special.polygamma(1, x)) carries up to ~5e-10 relative error in float64; the
cancellation in d ELBO / d w_1 amplifies it past the 1e-9 parity bar.  Both the value and the derivative therefore come from
the device functions of csrc/special.cuh through the C ABI (dpgp_polygamma), which the fused kernels use as well.
"""
import ctypes as C

import torch

from .. import _lib


# Test hook (same role as models.dp_gp_lvm.ENGINE_FACTORY): tests/test_distributed_cpu.py runs the host-side sharding logic
# on CPU tensors over gloo and injects a scipy-backed (digamma, trigamma) here.  The product never sets it.
POLYGAMMA_HOOK = None


def _polygamma(x, want_psi, want_tri):
    if POLYGAMMA_HOOK is not None:
        return POLYGAMMA_HOOK(x, want_psi, want_tri)
    if not x.is_cuda:
        raise RuntimeError("dp_gp_lvm_b200 needs CUDA tensors (no CPU fallback)")
    xc = x.detach().contiguous().to(torch.float64)
    psi = torch.empty_like(xc) if want_psi else None
    tri = torch.empty_like(xc) if want_tri else None
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(x.device):
        rc = _lib.lib().dpgp_polygamma(ptr(xc), ptr(psi), ptr(tri), xc.numel(), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    if rc != 0:
        raise _lib.DpgpError(rc, "dpgp_polygamma failed")
    return psi, tri


class _Digamma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return _polygamma(x, True, False)[0].view(x.shape)

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        return grad * _polygamma(x, False, True)[1].view(x.shape)


def digamma(x):
    return _Digamma.apply(x)
