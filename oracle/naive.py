"""TEST INFRASTRUCTURE ONLY (never imported by dp_gp_lvm_b200/) -- known-answer functions.

A restatement, in plain numpy/scipy loops, of the "naive" re-computations the reference's own unit
tests use as their expected values.  They cannot be imported from the reference because every module
under /root/reference/test/unittests imports tensorflow at the top, so the maths is restated here
(the element-by-element definitions, not the code) with the reference location of each definition:

  covariance_matrix  test/unittests/kernel_unittests.py:14-53
  psi_0              test/unittests/kernel_unittests.py:71-78
  psi_1              test/unittests/kernel_unittests.py:81-111
  psi_2              test/unittests/kernel_unittests.py:114-147
  free energy        test/unittests/bgplvm_unittests.py:17-52   (explicit inverse / determinant form)
  KL(q(X)||p(X))     test/unittests/bgplvm_unittests.py:123-135
  DP ELBO terms      test/unittests/dp_unittests.py:13-131
  log-normal prior   src/distributions/log_normal.py:24-39
  D-mode composition test/unittests/dpgplvm_unitttests.py:78-126

PARITY PIN: `oracle/run_reference_unittests.py` runs the reference's own copies of these functions
against the reference's own model code (over the TF shim) in this container; `tests/test_oracle.py`
checks this restatement against the fixtures written by `oracle/make_golden.py` from that run.
Pure Python loops: use only at unit-test sizes (N<=200, M<=75).
"""
import math

import numpy as np
from scipy.special import digamma, gammaln

JITTER = 1.0e-8      # src/utils/constants.py:96


def covariance_matrix(x0, gamma, alpha, beta, x1=None, include_noise=False, include_jitter=False):
    """k(x, z) = alpha * exp(-1/2 sum_q gamma_q (x_q - z_q)^2); noise/jitter only for k(x, x)."""
    square = x1 is None
    if square:
        x1 = x0
    n0, q = x0.shape
    n1 = x1.shape[0]
    k = np.empty((n0, n1))
    for i in range(n0):
        for j in range(n1):
            acc = 0.0
            for c in range(q):
                diff = x0[i, c] - x1[j, c]
                acc += gamma[c] * diff * diff
            k[i, j] = alpha * math.exp(-0.5 * acc)
    if square and include_noise:
        k = k + np.eye(n0) / beta
    if square and include_jitter:
        k = k + JITTER * np.eye(n0)
    return k


def psi_0(num_samples, alpha):
    return num_samples * alpha


def psi_1(x_mean, x_var, x_u, gamma, alpha):
    n, q = x_mean.shape
    m = x_u.shape[0]
    out = np.empty((n, m))
    for i in range(n):
        for k in range(m):
            lg = math.log(alpha)
            for j in range(q):
                den = gamma[j] * x_var[i, j] + 1.0
                diff = x_mean[i, j] - x_u[k, j]
                lg -= 0.5 * (math.log(den) + gamma[j] * diff * diff / den)
            out[i, k] = math.exp(lg)
    return out


def psi_2(x_mean, x_var, x_u, gamma, alpha):
    n, q = x_mean.shape
    m = x_u.shape[0]
    out = np.zeros((m, m))
    for i in range(n):
        den = [2.0 * gamma[j] * x_var[i, j] + 1.0 for j in range(q)]
        half_log_den = sum(0.5 * math.log(d) for d in den)
        for k1 in range(m):
            for k2 in range(k1, m):
                lg = 2.0 * math.log(alpha) - half_log_den
                for j in range(q):
                    zbar = 0.5 * (x_u[k1, j] + x_u[k2, j])
                    dz = x_u[k1, j] - x_u[k2, j]
                    dm = x_mean[i, j] - zbar
                    lg -= 0.25 * gamma[j] * dz * dz + gamma[j] * dm * dm / den[j]
                v = math.exp(lg)
                out[k1, k2] += v
                if k2 != k1:
                    out[k2, k1] += v
    return out


def free_energy(y, x_mean, x_var, x_u, gamma, alpha, beta):
    """Collapsed bound term F for ONE kernel (gamma, alpha, beta) and data y [N x d], inverse/det form."""
    n = x_mean.shape[0]
    d = y.shape[1]
    k_uu = covariance_matrix(x_u, gamma, alpha, beta, include_jitter=True)
    p0 = psi_0(n, alpha)
    p1 = psi_1(x_mean, x_var, x_u, gamma, alpha)
    p2 = psi_2(x_mean, x_var, x_u, gamma, alpha)
    sigma = beta * p2 + k_uu
    w = beta * np.eye(n) - beta * beta * p1 @ np.linalg.inv(sigma) @ p1.T
    _, logdet_k = np.linalg.slogdet(k_uu)
    _, logdet_s = np.linalg.slogdet(sigma)
    return (0.5 * n * d * math.log(beta) + 0.5 * d * logdet_k - 0.5 * n * d * math.log(2.0 * math.pi)
            - 0.5 * d * logdet_s - 0.5 * d * beta * p0
            + 0.5 * d * beta * np.trace(np.linalg.solve(k_uu, p2)) - 0.5 * np.trace(w @ (y @ y.T)))


def kl_qx_px(x_mean, x_var):
    n, q = x_mean.shape
    return 0.5 * (np.sum(x_mean ** 2) + np.sum(x_var) - np.sum(np.log(x_var)) - n * q)


def log_normal_log_pdf(x):
    x = np.asarray(x, dtype=np.float64)
    return -np.log(x) - 0.5 * (math.log(2.0 * math.pi) + np.log(x) ** 2)


# ---------------------------------------------------------------------------------------------- DP
def dp_elbo_terms(phi, gamma_1, gamma_2, w_1, w_2, s_1, s_2):
    d, t = phi.shape
    e_z = 0.0
    for i in range(d):
        for k in range(t - 1):
            tail = sum(phi[i, j] for j in range(k + 1, t))
            dg12 = digamma(gamma_1[k] + gamma_2[k])
            e_z += phi[i, k] * (digamma(gamma_1[k]) - dg12) + tail * (digamma(gamma_2[k]) - dg12)
    sum_b = sum(digamma(gamma_2[k]) - digamma(gamma_1[k] + gamma_2[k]) for k in range(t - 1))
    e_v = (t - 1.0) * (digamma(w_1) - math.log(w_2)) + (w_1 / w_2 - 1.0) * sum_b
    e_a = s_1 * math.log(s_2) - gammaln(s_1) + (s_1 - 1.0) * (digamma(w_1) - math.log(w_2)) - s_2 * w_1 / w_2
    h_z = -sum(phi[i, k] * math.log(phi[i, k]) for i in range(d) for k in range(t))
    h_v = 0.0
    for k in range(t - 1):
        a, b = gamma_1[k], gamma_2[k]
        h_v += (gammaln(a) + gammaln(b) - gammaln(a + b) - (a - 1.0) * digamma(a) - (b - 1.0) * digamma(b)
                + (a + b - 2.0) * digamma(a + b))
    h_a = w_1 - math.log(w_2) + gammaln(w_1) + (1.0 - w_1) * digamma(w_1)
    return float(e_z), float(e_v), float(e_a), float(h_z), float(h_v), float(h_a)


def dp_elbo(phi, gamma_1, gamma_2, w_1, w_2, s_1, s_2):
    return sum(dp_elbo_terms(phi, gamma_1, gamma_2, w_1, w_2, s_1, s_2))


# ------------------------------------------------------------------------------ D-mode composition
def dmode_objective(y, x_mean, x_var, x_u, phi, gamma_1, gamma_2, w_1, w_2, s_1, s_2,
                    gamma_atoms, alpha_atoms, beta_atoms):
    """objective = -(DP ELBO + sum_d F(y_d; phi_d-mixed hyper-parameters) - KL + hyper-prior)."""
    gam = phi @ gamma_atoms                 # [D x Q]   src/models/dp_gp_lvm.py:100
    alp = (phi @ alpha_atoms).reshape(-1)   # [D]       :101
    bet = (phi @ beta_atoms).reshape(-1)    # [D]       :102
    gp = -kl_qx_px(x_mean, x_var)
    for d in range(y.shape[1]):
        gp += free_energy(y[:, d:d + 1], x_mean, x_var, x_u, gam[d], alp[d], bet[d])
    prior = (np.sum(log_normal_log_pdf(gamma_atoms)) + np.sum(log_normal_log_pdf(alpha_atoms))
             + np.sum(log_normal_log_pdf(beta_atoms)))
    return -(dp_elbo(phi, gamma_1, gamma_2, w_1, w_2, s_1, s_2) + gp + prior)
