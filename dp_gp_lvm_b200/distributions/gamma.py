"""Gamma(shape alpha, rate beta) entropy (reference src/distributions/gamma.py:8-17)."""
import torch

from ..utils.special import digamma


def entropy(alpha, beta):
    return alpha - torch.log(beta) + torch.lgamma(alpha) + (1.0 - alpha) * digamma(alpha)
