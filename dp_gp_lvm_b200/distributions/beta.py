"""Beta entropy (reference src/distributions/beta.py:8-19)."""
import torch

from ..utils.special import digamma


def entropy(alpha, beta):
    total = alpha + beta
    return (torch.lgamma(alpha) + torch.lgamma(beta) - torch.lgamma(total) - (alpha - 1.0) * digamma(alpha)
            - (beta - 1.0) * digamma(beta) + (total - 2.0) * digamma(total))
