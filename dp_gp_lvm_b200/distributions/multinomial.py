"""Categorical entropy -sum p log p over the last axis (reference src/distributions/multinomial.py:8-16)."""
import torch


def entropy(probs):
    return -(probs * torch.log(probs)).sum(-1)
