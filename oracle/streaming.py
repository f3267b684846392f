"""TEST INFRASTRUCTURE ONLY (never imported by dp_gp_lvm_b200/) -- O(N)-memory oracle.

The reference graph materialises a [B,N,M,M,Q] tensor for psi_2 (src/kernels/rbf_kernel.py:194) and,
in D-mode, a [D,N,N] one (src/models/dp_gp_lvm.py:134); neither exists for the Frey-shaped and N=1M
configurations.  Every N-dependent quantity enters the objective only through sums over n, so this
module evaluates THE SAME objective from streamed sufficient statistics:

    Psi2_b = sum_n psi2_n,   P_b = Psi1_b^T Y  (T-mode)  /  p_d = Psi1_d^T y_d  (D-mode),
    yy_d = sum_n y_nd^2,     KL sums,

with  ||L_A^-1 L^-1 p||^2 = (beta C y)^2-term of dp_gp_lvm.py:657-658 / :145.  The M x M chain keeps the
reference's operation order (dp_gp_lvm.py:618-635).  Gradients: two chunked passes with torch autograd
(pass 1 accumulates the statistics, the bound is differentiated w.r.t. them, pass 2 re-computes each
chunk's statistics under autograd and back-propagates the cotangents), which is exact.

tests/test_oracle.py checks streaming == literal (<= 1e-12 relative) at every golden shape, so the
large-N values this module produces inherit the pin of oracle/literal.py.
"""
import math

import numpy as np
import torch

from oracle import literal as L

LOG_2PI = L.LOG_2PI


def chunk_stats(x_u, x_mean, x_var, y, gamma, alpha, mode):
    """Statistics of one chunk of rows.  Returns (Psi2 [B,M,M], P ([B,M,D] or [D,M]))."""
    b = gamma.shape[0]
    m = x_u.shape[0]
    # psi1 [B,n,M]
    den1 = gamma[:, None, :] * x_var[None] + 1.0
    diff = x_mean[:, None, :] - x_u[None]                                   # [n,M,Q]
    e1 = torch.einsum("bnq,nmq->bnm", gamma[:, None, :] / den1, diff ** 2)
    p1 = torch.exp(torch.log(alpha)[:, :, None] - 0.5 * (e1 + torch.log(den1).sum(-1)[:, :, None]))
    if mode == "t":
        p = torch.einsum("bnm,nd->bmd", p1, y)
    else:
        p = torch.einsum("dnm,nd->dm", p1, y)
    # psi2 [B,M,M], looping over clusters to bound memory: [n,M,M,Q]
    zbar = 0.5 * (x_u[:, None, :] + x_u[None, :, :])
    dz2 = (x_u[:, None, :] - x_u[None, :, :]) ** 2
    p2 = []
    for k in range(b):
        g = gamma[k]
        den = 2.0 * g * x_var + 1.0                                         # [n,Q]
        num = ((x_mean[:, None, None, :] - zbar[None]) ** 2 * (g / den)[:, None, None, :]).sum(-1)   # [n,M,M]
        lg = 2.0 * torch.log(alpha[k]) - 0.5 * torch.log(den).sum(-1)[:, None, None] - 0.25 * (dz2 * g).sum(-1)[None] - num
        p2.append(torch.exp(lg).sum(0))
    return torch.stack(p2), p


def bound_from_stats(n, d, p2, p, yy, kl_sum_mu2, kl_sum_s, x_u, gamma, alpha, beta, phi, mode):
    """f_hat - KL from the sufficient statistics (T-mode: SURVEY Appendix B.1; D-mode likewise)."""
    l, h, la = L._chain(L.k_uu(x_u, gamma, alpha), p2, beta)
    tr_h = torch.diagonal(h, dim1=1, dim2=2).sum(-1, keepdim=True)
    logdet_la = torch.log(torch.diagonal(la, dim1=1, dim2=2)).sum(-1, keepdim=True)
    kl = 0.5 * (kl_sum_mu2 + kl_sum_s - float(n) * x_u.shape[1])
    if mode == "t":
        c = torch.linalg.solve_triangular(la, torch.linalg.solve_triangular(l, p, upper=False), upper=False)  # [T,M,D]
        qf = (c ** 2).sum(1)                                                 # [T,D]
        phi_td = phi.transpose(0, 1)
        f_hat = (-0.5 * n * d * LOG_2PI
                 + (phi_td * (0.5 * (n * torch.log(beta) + beta * (tr_h - alpha * n)) - logdet_la)).sum()
                 - 0.5 * (phi_td * beta * yy[None]).sum() + 0.5 * (phi_td * beta ** 2 * qf).sum())
    else:
        c = torch.linalg.solve_triangular(la, torch.linalg.solve_triangular(l, p[:, :, None], upper=False),
                                          upper=False)[:, :, 0]            # [D,M]
        qf = (c ** 2).sum(1, keepdim=True)                                   # [D,1]
        f_hat = (0.5 * n * (torch.log(beta).sum() - d * LOG_2PI) - logdet_la.sum()
                 + 0.5 * (beta * (tr_h - alpha * n)).sum() - 0.5 * (beta * yy[:, None]).sum()
                 + 0.5 * (beta ** 2 * qf).sum())
    return f_hat - kl


def value_and_grad(y, params_np, mode="t", alpha_prior=(1.0, 1.0), mask_size=1, chunk=2048, want_grad=True):
    """Chunked evaluation of objective (and gradients w.r.t. the raw parameters)."""
    y = torch.as_tensor(np.asarray(y, dtype=np.float64))
    n, d = y.shape
    leaf = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in params_np.items()}

    def hyper(c):
        if mode == "t":
            return c["gamma_atoms"], c["alpha_atoms"], c["beta_atoms"]
        return c["phi"] @ c["gamma_atoms"], c["phi"] @ c["alpha_atoms"], c["phi"] @ c["beta_atoms"]

    # pass 1: statistics, no graph
    with torch.no_grad():
        c0 = L.constrained(leaf, d, mask_size)
        gam0, alp0, _ = hyper(c0)
        p2 = None; p = None
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            a, b_ = chunk_stats(c0["x_u"], c0["x_mean"][s:e], c0["x_var"][s:e], y[s:e], gam0, alp0, mode)
            p2 = a if p2 is None else p2 + a
            p = b_ if p is None else p + b_
        yy = (y ** 2).sum(0)
    # the small part under autograd, statistics as leaves
    p2_l = p2.clone().requires_grad_(True)
    p_l = p.clone().requires_grad_(True)
    small = {k: v for k, v in leaf.items()}
    c = L.constrained(small, d, mask_size)
    gam, alp, bet = hyper(c)
    dp = L.dp_objective(c["phi"], c["g1"], c["g2"], c["w1"], c["w2"], float(alpha_prior[0]), float(alpha_prior[1]))
    prior = (L.log_normal_log_pdf(c["gamma_atoms"]).sum() + L.log_normal_log_pdf(c["alpha_atoms"]).sum()
             + L.log_normal_log_pdf(c["beta_atoms"]).sum())
    gp = bound_from_stats(n, d, p2_l, p_l, yy, (c["x_mean"] ** 2).sum(), (c["x_var"] - torch.log(c["x_var"])).sum(),
                          c["x_u"], gam, alp, bet, c["phi"], mode)
    obj = dp - gp - prior
    if not want_grad:
        return float(obj.detach()), None
    names = list(L.PARAM_ORDER)
    g_all = torch.autograd.grad(obj, [leaf[k] for k in names] + [p2_l, p_l], allow_unused=True)
    grads = {k: (torch.zeros_like(leaf[k]) if g is None else g.clone()) for k, g in zip(names, g_all[:len(names)])}
    g_p2, g_p = g_all[-2], g_all[-1]
    # pass 2: chunk statistics under autograd, cotangents g_p2 / g_p
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        lf = {k: leaf[k].detach().clone().requires_grad_(True) for k in names}
        cc = L.constrained(lf, d, mask_size)
        gam_c, alp_c, _ = hyper(cc)
        a, b_ = chunk_stats(cc["x_u"], cc["x_mean"][s:e], cc["x_var"][s:e], y[s:e], gam_c, alp_c, mode)
        sur = (a * g_p2).sum() + (b_ * g_p).sum()
        gs = torch.autograd.grad(sur, [lf[k] for k in names], allow_unused=True)
        for k, g in zip(names, gs):
            if g is not None:
                grads[k] += g
    return float(obj.detach()), {k: v.numpy().copy() for k, v in grads.items()}


def gp_value_and_grad(y, mu, s, z, gamma, alpha, beta, phi, mode="t", chunk=4096):
    """Stage-level oracle for dpgp_stats_fwd + dpgp_bound + dpgp_stats_bwd: value of gp = f_hat - KL, the
    packed statistics, and the gradients of gp w.r.t. (mu, s, z, gamma [B,Q], alpha [B,1], beta [B,1], phi).
    numpy in / numpy out; single autograd graph (use at sizes where [n,M,M,Q] per cluster fits)."""
    t = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=True)
    y = torch.as_tensor(np.asarray(y, dtype=np.float64))
    mu, s, z, gamma, alpha, beta = t(mu), t(s), t(z), t(gamma), t(np.reshape(alpha, (-1, 1))), t(np.reshape(beta, (-1, 1)))
    phi_t = t(phi) if phi is not None else None
    n, d = y.shape
    p2 = None; p = None
    for a in range(0, n, chunk):
        e = min(n, a + chunk)
        x2, x1 = chunk_stats(z, mu[a:e], s[a:e], y[a:e], gamma, alpha, mode)
        p2 = x2 if p2 is None else p2 + x2
        p = x1 if p is None else p + x1
    yy = (y ** 2).sum(0)
    gp = bound_from_stats(n, d, p2, p, yy, (mu ** 2).sum(), (s - torch.log(s)).sum(), z, gamma, alpha, beta, phi_t, mode)
    leaves = [mu, s, z, gamma, alpha, beta] + ([phi_t] if phi_t is not None else [])
    grads = torch.autograd.grad(gp, leaves)
    names = ["mu", "s", "z", "gamma", "alpha", "beta"] + (["phi"] if phi_t is not None else [])
    out = {k: g.numpy().copy() for k, g in zip(names, grads)}
    stats = dict(psi2=p2.detach().numpy(), p=p.detach().numpy(), yy=yy.numpy(),
                 kl=np.array([float((mu ** 2).sum()), float((s - torch.log(s)).sum())]))
    return float(gp.detach()), stats, out
