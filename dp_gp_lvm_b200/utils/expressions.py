"""Initialisation helpers (reference src/utils/expressions.py).

`principal_component_analysis` returns the same quantity as the reference (:47-76) -- the leading
eigenvectors of Y Y^T scaled by the mean of their column standard deviations -- but obtains them from the
D x D matrix Y^T Y (u_i = Y v_i / sigma_i) instead of an N x N eigen-problem, which is what makes N = 1M
initialisable.  Eigenvector signs are arbitrary in both (ARPACK's start vector is unseeded)."""
import numpy as np


def principal_component_analysis(x, num_latent_dimensions):
    assert isinstance(x, np.ndarray)
    assert x.ndim == 2
    n, d = x.shape
    assert 0 < num_latent_dimensions < min(n, d), \
        'Number of latent dimensions must be greater than zero and less than the minimum of the number of ' \
        'observations and the number of observed dimensions.'
    if n <= d:
        w, v = np.linalg.eigh(x @ x.T)
        x_0 = v[:, ::-1][:, :num_latent_dimensions]
    else:
        w, v = np.linalg.eigh(x.T @ x)
        idx = np.argsort(w)[::-1][:num_latent_dimensions]
        x_0 = (x @ v[:, idx]) / np.sqrt(np.maximum(w[idx], 1e-300))
    x_0 = x_0 / np.mean(x_0.std(axis=0, ddof=1))
    return x_0


def print_and_log(*args, **kwargs):
    print(*args, **kwargs)
