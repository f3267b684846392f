"""KL(q(X) || N(0, I)) (reference src/models/expressions/gp_expressions.py:10-24).

Stand-alone torch version for API parity; the bound path fuses the two sums into its statistics pass
(colsum_kernel) and never calls this."""
import torch


def calculate_kl_divergence_standard_prior(x_mean, x_covar):
    n, q = x_mean.shape
    diag = torch.diagonal(x_covar, dim1=-2, dim2=-1) if x_covar.dim() == 3 else x_covar
    return 0.5 * ((x_mean ** 2).sum() + (diag - torch.log(diag)).sum() - float(n * q))
