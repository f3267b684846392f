"""Beta entropy (reference src/distributions/beta.py:8-19)."""
import torch


def entropy(alpha, beta):
    total = alpha + beta
    return (torch.lgamma(alpha) + torch.lgamma(beta) - torch.lgamma(total) - (alpha - 1.0) * torch.digamma(alpha)
            - (beta - 1.0) * torch.digamma(beta) + (total - 2.0) * torch.digamma(total))
