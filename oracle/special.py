"""TEST INFRASTRUCTURE ONLY -- digamma with an accurate derivative for the CPU oracle.

torch.digamma is accurate to ~1e-15 in float64, but its autograd derivative (torch.special.polygamma(1, x)) cuts the
asymptotic series after the x^-7 term and carries up to ~5e-10 relative error; the cancellation in d ELBO / d w_1 of the
DP objective (src/models/dirichlet_process.py:68-77) amplifies that to 1e-8 .. 4e-7 in the gradient of w1_raw (measured
at the BASELINE shapes, round 2).  TensorFlow's own polygamma kernel (Eigen zeta-function series) does not have this
defect, so an oracle that is to stand in for `tf.gradients` needs an accurate trigamma: scipy.special.polygamma(1, x)
(Hurwitz zeta, ~1e-16) is used in the backward pass here.  Forward values are unchanged (torch.digamma).
"""
import numpy as np
import torch
from scipy.special import polygamma


class _Digamma(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.digamma(x)

    @staticmethod
    def backward(ctx, grad):
        (x,) = ctx.saved_tensors
        tri = torch.as_tensor(polygamma(1, x.detach().cpu().numpy().astype(np.float64)), dtype=x.dtype).reshape(x.shape)
        return grad * tri


def digamma(x):
    return _Digamma.apply(x)
