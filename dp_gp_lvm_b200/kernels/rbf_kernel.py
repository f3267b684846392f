"""RBF kernels (reference src/kernels/rbf_kernel.py).  Only the ARD form exists in the reference
(`k_rbf` and `k_mahalanobis_rbf` raise NotImplementedError there, :13-23 and :206-218; same here).

`k_ard_rbf(gamma [B,Q], alpha [B,1], beta [B,1])` returns the reference's `Kernel` object whose closures
run the CUDA kernels behind include/dpgp.h for a batch of B kernels:
    covariance_matrix  -> dpgp_covariance   (rbf_kernel.py:58-93)
    covariance_diag    -> closed form       (rbf_kernel.py:96-116)
    psi_0              -> alpha * N         (rbf_kernel.py:119-132)
    psi_1              -> dpgp_psi1         (rbf_kernel.py:135-161)
    psi_2              -> dpgp_stats_fwd    (rbf_kernel.py:164-199)
These stand-alone statistics are forward-only (they are API surface and test hooks); the differentiable
path is the fused bound in models/dp_gp_lvm.py, which never materialises Psi1.
"""
import numpy as np
import torch

from .. import engine as _engine
from ..distributions.log_normal import log_pdf as log_normal_log_pdf
from ..utils.constants import GP_DEFAULT_JITTER
from .interfaces.kernel import Kernel, KernelHyperparameters, _value


def k_rbf(gamma, alpha, beta):
    raise NotImplementedError


def k_mahalanobis_rbf(weights, gamma, alpha, beta):
    raise NotImplementedError


def _dev_tensor(x, device):
    if isinstance(x, torch.Tensor):
        return x.detach().to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.asarray(x, dtype=np.float64), device=device).contiguous()


def _diag_part(cov):
    """latent_input_covariance is [N,Q,Q] in the reference and only its diagonal is read (:151, :182)."""
    return torch.diagonal(cov, dim1=-2, dim2=-1).contiguous() if cov.dim() == 3 else cov


def k_ard_rbf(gamma, alpha, beta, device=None):
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    engines = {}

    def hp():
        g = _dev_tensor(_value(gamma), device)
        a = _dev_tensor(_value(alpha), device).reshape(-1)
        b = _dev_tensor(_value(beta), device).reshape(-1)
        assert g.dim() == 2 and a.numel() == g.shape[0] and b.numel() == g.shape[0], \
            'gamma must be [B x Q], alpha and beta [B x 1].'
        return g, a, b

    def eng(n, d, q, m, b):
        key = (n, d, q, m, b)
        if key not in engines:
            engines[key] = _engine.BoundEngine(n, d, q, m, b, _engine.MODE_T, device=device)
        return engines[key]

    hyperparameters_dict = {KernelHyperparameters.ARD_WEIGHTS: gamma,
                            KernelHyperparameters.SIGNAL_VARIANCE: alpha,
                            KernelHyperparameters.NOISE_PRECISION: beta}
    hyperpriors_dict = {KernelHyperparameters.ARD_WEIGHTS: log_normal_log_pdf,
                        KernelHyperparameters.SIGNAL_VARIANCE: log_normal_log_pdf,
                        KernelHyperparameters.NOISE_PRECISION: log_normal_log_pdf}

    def covariance_matrix_func(input_0, input_1=None, include_noise=False, include_jitter=False):
        g, a, b = hp()
        x0 = _dev_tensor(input_0, device)
        x1 = None if input_1 is None else _dev_tensor(input_1, device)
        e = eng(1, 1, g.shape[1], 1, g.shape[0])
        return e.covariance(x0, x1, g, a, b, include_noise=include_noise, include_jitter=include_jitter)

    def covariance_diagonal_func(input_0, include_noise=False, include_jitter=False):
        g, a, b = hp()
        n = input_0.shape[0]
        k = a[:, None] * torch.ones((1, n), dtype=torch.float64, device=device)
        if include_noise:
            k = k + 1.0 / b[:, None]
        if include_jitter:
            k = k + GP_DEFAULT_JITTER
        return k

    def calculate_psi_0(inducing_input, latent_input_mean, latent_input_covariance):
        g, a, b = hp()
        return a[:, None] * float(latent_input_mean.shape[0])

    def calculate_psi_1(inducing_input, latent_input_mean, latent_input_covariance):
        g, a, b = hp()
        z = _dev_tensor(inducing_input, device); mu = _dev_tensor(latent_input_mean, device)
        s = _diag_part(_dev_tensor(latent_input_covariance, device))
        e = eng(mu.shape[0], 1, g.shape[1], z.shape[0], g.shape[0])
        return e.psi1(mu, s, z, g, a)

    def calculate_psi_2(inducing_input, latent_input_mean, latent_input_covariance):
        g, a, b = hp()
        z = _dev_tensor(inducing_input, device); mu = _dev_tensor(latent_input_mean, device)
        s = _diag_part(_dev_tensor(latent_input_covariance, device))
        e = eng(mu.shape[0], 1, g.shape[1], z.shape[0], g.shape[0])
        y0 = torch.zeros((mu.shape[0], 1), dtype=torch.float64, device=device)
        stats = e.stats_fwd(mu, s, y0, z, g, a)
        return e.split_stats(stats)[0].clone()

    return Kernel(covar_matrix_func=covariance_matrix_func, covar_diag_func=covariance_diagonal_func,
                  hyperparameter_dict=hyperparameters_dict, hyperprior_func_dict=hyperpriors_dict,
                  psi_0_func=calculate_psi_0, psi_1_func=calculate_psi_1, psi_2_func=calculate_psi_2)
