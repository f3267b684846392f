"""Univariate log-normal log-pdf (reference src/distributions/log_normal.py:24-39); default mean 0, var 1."""
import math

import torch


def log_pdf(x, mean=None, var=None):
    if mean is None:
        mean = torch.zeros_like(x)
    if var is None:
        var = torch.ones_like(x)
    return -torch.log(x) - 0.5 * (torch.log(2.0 * math.pi * var) + (torch.log(x) - mean) ** 2 / var)


def pdf(x, mean=None, var=None):
    return torch.exp(log_pdf(x, mean, var))


def entropy(mean=None, var=None):
    raise NotImplementedError
