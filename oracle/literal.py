"""TEST INFRASTRUCTURE ONLY (never imported by dp_gp_lvm_b200/) -- the CPU float64 oracle.

A restatement in PyTorch-CPU float64 of the reference's DP-GP-LVM objective with the SAME op sequence
and the same materialised intermediates as the TensorFlow graph, so that torch autograd yields the
gradients `tf.train.AdamOptimizer.minimize` would see.  Reference locations:

  softplus / positive variables      src/utils/types.py:40-72
  RBF-ARD K_uu (+1e-8 I)             src/kernels/rbf_kernel.py:58-93
  psi_0 / psi_1 / psi_2              src/kernels/rbf_kernel.py:119-132 / 135-161 / 164-199
  DP objective                       src/models/dirichlet_process.py:39-88
  entropies, log-normal prior        src/distributions/{beta,gamma,multinomial,log_normal}.py
  KL(q(X)||N(0,I))                   src/models/expressions/gp_expressions.py:10-24
  D-mode bound (`dp_gp_lvm`)         src/models/dp_gp_lvm.py:84-154
  T-mode bound (`dp_gp_lvm_t`)       src/models/dp_gp_lvm.py:582-676

PARITY PIN: tests/golden/*.npz hold objective values AND gradients produced by the reference's own
source (executed over oracle/tf_shim by oracle/make_golden.py); tests/test_oracle.py checks this module
against them, and against oracle/naive.py (the reference's known-answer functions).  TensorFlow's own
numerical kernels are not available here, so agreement with TF itself is pinned only through the
reference's unit-test tolerance (rtol 1e-7) -- see DESIGN.md "Oracle".

The parameter dictionary everywhere in this repository (raw = unconstrained, as the reference's
tf.Variables, in the reference's creation order):
  x_mean [N,Q], x_var_raw [N,Q], x_u [M,Q], phi_logits [D/mask,T], gamma1_raw [T-1], gamma2_raw [T-1],
  w1_raw [], w2_raw [], gamma_atoms_raw [T,Q], alpha_atoms_raw [T,1], beta_atoms_raw [T,1]
"""
import math

import numpy as np
import torch

from oracle.special import digamma as _digamma          # accurate trigamma in the backward pass

JITTER = 1.0e-8
PARAM_ORDER = ("x_mean", "x_var_raw", "x_u", "phi_logits", "gamma1_raw", "gamma2_raw", "w1_raw", "w2_raw",
               "gamma_atoms_raw", "alpha_atoms_raw", "beta_atoms_raw")
LOG_2PI = math.log(2.0 * math.pi)


def softplus(x):
    return torch.nn.functional.softplus(x, beta=1.0, threshold=1.0e9)


def inv_softplus(v):
    return np.log(np.expm1(v))


# ------------------------------------------------------------------------------------------ kernel
def k_uu(x_u, gamma, alpha, jitter=True):
    """[B,M,M]; gamma [B,Q], alpha [B,1].  Same expansion as rbf_kernel.py:70-78."""
    sg = torch.sqrt(gamma)[:, None, :] * x_u[None]                 # [B,M,Q]
    sq = -0.5 * (sg * sg).sum(-1, keepdim=True)                    # [B,M,1]
    k = alpha[:, :, None] * torch.exp(sq + sq.transpose(1, 2) + sg @ sg.transpose(1, 2))
    if jitter:
        k = k + JITTER * torch.eye(x_u.shape[0], dtype=x_u.dtype)
    return k


def psi_0(n, alpha):
    return alpha * float(n)                                        # [B,1]


def psi_1(x_u, x_mean, x_var, gamma, alpha):
    """[B,N,M] materialising the [B,N,M,Q] intermediate exactly as rbf_kernel.py:155-161."""
    den = gamma[:, None, :] * x_var[None] + 1.0                                      # [B,N,Q]
    num = gamma[:, None, None, :] * (x_mean[:, None, :] - x_u[None]) ** 2            # [B,N,M,Q]
    lg = torch.log(alpha)[:, :, None] - 0.5 * (num / den[:, :, None, :] + torch.log(den)[:, :, None, :]).sum(-1)
    return torch.exp(lg)


def psi_2(x_u, x_mean, x_var, gamma, alpha):
    """[B,M,M] materialising [B,N,M,M,Q] exactly as rbf_kernel.py:189-199 (small shapes only)."""
    g = gamma[:, None, None, None, :]
    zbar = 0.5 * (x_u[:, None, :] + x_u[None, :, :])                                 # [M,M,Q]
    t1 = 0.25 * g * ((x_u[:, None, :] - x_u[None, :, :]) ** 2)[None, None]           # [B,1,M,M,Q]
    den = 2.0 * g * x_var[None, :, None, None, :] + 1.0                              # [B,N,1,1,Q]
    num = g * (x_mean[None, :, None, None, :] - zbar[None, None]) ** 2               # [B,N,M,M,Q]
    lg = 2.0 * torch.log(alpha)[:, :, None, None] - (0.5 * torch.log(den) + t1 + num / den).sum(-1)
    return torch.exp(lg).sum(1)


# ---------------------------------------------------------------------------------------------- DP
def phi_from_logits(logits, num_dims, mask_size):
    sm = torch.softmax(logits, dim=-1)
    if mask_size == 1:
        return sm
    depth = num_dims // mask_size
    idx = torch.as_tensor(np.repeat(np.arange(depth), mask_size))
    return torch.nn.functional.one_hot(idx, depth).to(sm.dtype) @ sm                # dirichlet_process.py:44-51


def dp_objective(phi, g1, g2, w1, w2, s1, s2):
    """-(ELBO) of the truncated stick-breaking DP, dirichlet_process.py:64-88."""
    t = phi.shape[1]
    dg1, dg2, dg12 = _digamma(g1), _digamma(g2), _digamma(g1 + g2)
    tail = torch.flip(torch.cumsum(torch.flip(phi, [1]), 1), [1]) - phi              # exclusive reverse cumsum
    e_z = (phi[:, :-1] * (dg1 - dg12) + tail[:, :-1] * (dg2 - dg12)).sum()
    e_v = (t - 1.0) * (_digamma(w1) - torch.log(w2)) + (w1 / w2 - 1.0) * (dg2 - dg12).sum()
    e_a = s1 * math.log(s2) - math.lgamma(s1) + (s1 - 1.0) * (_digamma(w1) - torch.log(w2)) - s2 * w1 / w2
    h_z = -(phi * torch.log(phi)).sum()
    h_v = (torch.lgamma(g1) + torch.lgamma(g2) - torch.lgamma(g1 + g2) - (g1 - 1.0) * dg1 - (g2 - 1.0) * dg2
           + (g1 + g2 - 2.0) * dg12).sum()
    h_a = w1 - torch.log(w2) + torch.lgamma(w1) + (1.0 - w1) * _digamma(w1)
    return -(e_z + e_v + e_a + h_z + h_v + h_a)


def log_normal_log_pdf(x):
    return -torch.log(x) - 0.5 * (LOG_2PI + torch.log(x) ** 2)


def kl_standard_prior(x_mean, x_var):
    n, q = x_mean.shape
    return 0.5 * ((x_mean ** 2).sum() + (x_var - torch.log(x_var)).sum() - float(n * q))


# ------------------------------------------------------------------------------------- the bounds
def _chain(k, p2, beta):
    """L=chol(K); H=L^-1 Psi2 L^-T; A=beta H + I; L_A=chol(A)   (dp_gp_lvm.py:113-127 / :618-633)."""
    l = torch.linalg.cholesky(k)
    h = torch.linalg.solve_triangular(l, torch.linalg.solve_triangular(l, p2, upper=False).transpose(1, 2),
                                      upper=False).transpose(1, 2)
    a = beta[:, :, None] * h + torch.eye(k.shape[-1], dtype=k.dtype)
    la = torch.linalg.cholesky(a)
    return l, h, la


def constrained(params, num_dims, mask_size=1):
    """raw parameter dict -> constrained quantities (all torch float64)."""
    out = dict(
        x_mean=params["x_mean"], x_var=softplus(params["x_var_raw"]), x_u=params["x_u"],
        phi=phi_from_logits(params["phi_logits"], num_dims, mask_size),
        g1=softplus(params["gamma1_raw"]), g2=softplus(params["gamma2_raw"]),
        w1=softplus(params["w1_raw"]), w2=softplus(params["w2_raw"]),
        gamma_atoms=softplus(params["gamma_atoms_raw"]), alpha_atoms=softplus(params["alpha_atoms_raw"]),
        beta_atoms=softplus(params["beta_atoms_raw"]))
    return out


def _common(y, params, alpha_prior, mask_size):
    c = constrained(params, y.shape[1], mask_size)
    dp = dp_objective(c["phi"], c["g1"], c["g2"], c["w1"], c["w2"], float(alpha_prior[0]), float(alpha_prior[1]))
    prior = (log_normal_log_pdf(c["gamma_atoms"]).sum() + log_normal_log_pdf(c["alpha_atoms"]).sum()
             + log_normal_log_pdf(c["beta_atoms"]).sum())
    kl = kl_standard_prior(c["x_mean"], c["x_var"])
    return c, dp, prior, kl


def objective_t(y, params, alpha_prior=(1.0, 1.0), mask_size=1):
    """`dp_gp_lvm_t(...).objective`, src/models/dp_gp_lvm.py:582-676 (one kernel per DP atom)."""
    n, d = y.shape
    c, dp, prior, kl = _common(y, params, alpha_prior, mask_size)
    gam, alp, bet = c["gamma_atoms"], c["alpha_atoms"], c["beta_atoms"]
    phi_td = c["phi"].transpose(0, 1)
    p0 = psi_0(n, alp)
    p1 = psi_1(c["x_u"], c["x_mean"], c["x_var"], gam, alp)
    p2 = psi_2(c["x_u"], c["x_mean"], c["x_var"], gam, alp)
    l, h, la = _chain(k_uu(c["x_u"], gam, alp), p2, bet)
    tr_h = torch.diagonal(h, dim1=1, dim2=2).sum(-1, keepdim=True)                      # [T,1]
    logdet_la = torch.log(torch.diagonal(la, dim1=1, dim2=2)).sum(-1, keepdim=True)     # [T,1]
    cmat = torch.linalg.solve_triangular(la, torch.linalg.solve_triangular(l, p1.transpose(1, 2), upper=False),
                                         upper=False)                                   # [T,M,N]
    g = bet[:, :, None] * cmat
    recon = (phi_td * ((g @ y[None].expand(g.shape[0], n, d)) ** 2).sum(1)).sum()
    yy = (y ** 2).sum(0, keepdim=True)                                                  # [1,D]
    f_hat = (-0.5 * n * d * LOG_2PI
             + (phi_td * (0.5 * (n * torch.log(bet) + bet * (tr_h - p0)) - logdet_la)).sum()
             - 0.5 * (phi_td * bet * yy).sum() + 0.5 * recon)
    return dp - (f_hat - kl) - prior


def objective_d(y, params, alpha_prior=(1.0, 1.0), mask_size=1):
    """`dp_gp_lvm(...).objective`, src/models/dp_gp_lvm.py:84-154 (hyper-parameters mixed by phi, B = D)."""
    n, d = y.shape
    c, dp, prior, kl = _common(y, params, alpha_prior, mask_size)
    gam = c["phi"] @ c["gamma_atoms"]
    alp = c["phi"] @ c["alpha_atoms"]
    bet = c["phi"] @ c["beta_atoms"]
    p0 = psi_0(n, alp)
    p1 = psi_1(c["x_u"], c["x_mean"], c["x_var"], gam, alp)
    p2 = psi_2(c["x_u"], c["x_mean"], c["x_var"], gam, alp)
    l, h, la = _chain(k_uu(c["x_u"], gam, alp), p2, bet)
    logdet_la = torch.log(torch.diagonal(la, dim1=1, dim2=2)).sum()
    cmat = torch.linalg.solve_triangular(la, torch.linalg.solve_triangular(l, p1.transpose(1, 2), upper=False),
                                         upper=False)                                   # [D,M,N]
    ctc = cmat.transpose(1, 2) @ cmat                                                   # [D,N,N]
    yb = (y.transpose(0, 1) * bet)[:, None, :]                                          # [D,1,N]
    tr_h = torch.diagonal(h, dim1=1, dim2=2).sum(-1, keepdim=True)
    f_hat = (0.5 * n * (torch.log(bet).sum() - d * LOG_2PI) - logdet_la
             + 0.5 * (bet * (tr_h - p0)).sum()
             - 0.5 * (bet * torch.diagonal(y.transpose(0, 1) @ y)[:, None]).sum()
             + 0.5 * (yb @ ctc @ yb.transpose(1, 2)).sum())
    return dp - (f_hat - kl) - prior


def value_and_grad(fn, y, params_np, alpha_prior=(1.0, 1.0), mask_size=1):
    """numpy in / numpy out convenience: returns (objective, {name: gradient})."""
    y_t = torch.as_tensor(np.asarray(y, dtype=np.float64))
    p = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in params_np.items()}
    obj = fn(y_t, p, alpha_prior, mask_size)
    grads = torch.autograd.grad(obj, [p[k] for k in PARAM_ORDER])
    return float(obj.detach()), {k: g.numpy().copy() for k, g in zip(PARAM_ORDER, grads)}


def random_params(rng, n, d, q, m, t, mask_size=1, z_from_x=False, spread=0.3):
    """SURVEY.md 8(d) evaluation point: clusters differ so T-mode and D-mode do not coincide."""
    x_mean = rng.standard_normal((n, q))
    if z_from_x:
        x_u = x_mean[rng.permutation(n)[:m]] + 0.01 * rng.standard_normal((m, q))
    else:
        x_u = rng.standard_normal((m, q))
    base = inv_softplus(1.0)
    return dict(
        x_mean=x_mean, x_var_raw=base + 0.1 * rng.standard_normal((n, q)), x_u=x_u,
        phi_logits=rng.standard_normal((d // mask_size, t)),
        gamma1_raw=rng.standard_normal(t - 1), gamma2_raw=rng.standard_normal(t - 1),
        w1_raw=np.array(base), w2_raw=np.array(base),
        gamma_atoms_raw=base + spread * rng.standard_normal((t, q)),
        alpha_atoms_raw=base + spread * rng.standard_normal((t, 1)),
        beta_atoms_raw=base + spread * rng.standard_normal((t, 1)))
