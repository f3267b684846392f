"""float64 everywhere and softplus-positive variables (reference src/utils/types.py:13-72).

TensorFlow variables become torch leaf tensors (requires_grad=True) on the model's CUDA device; the
"positive node" softplus(raw) is re-evaluated whenever it is read, which is what a TF-1 graph does on
every session.run."""
import numpy as np
import torch

TORCH_DTYPE = torch.float64
NP_DTYPE = np.float64


class PositiveVariable:
    """`create_positive_variable`: positive = softplus(raw), raw_0 = log(exp(v_0) - 1) (types.py:40-57)."""

    def __init__(self, raw):
        self.raw = raw

    @property
    def value(self):
        return torch.nn.functional.softplus(self.raw, beta=1.0, threshold=1.0e9)


def create_positive_variable(initial_value, shape=None, is_trainable=True, device=None):
    assert initial_value > 0, 'Initial value must be positive.'
    init = np.log(np.exp(initial_value) - 1.0) * np.ones(shape=() if shape is None else shape, dtype=NP_DTYPE)
    raw = torch.tensor(init, dtype=TORCH_DTYPE, device=device, requires_grad=True)
    raw.is_trainable = is_trainable
    return PositiveVariable(raw)


def create_random_positive_variable(shape, is_trainable=True, device=None):
    """softplus(N(0,1)) drawn from numpy's global RNG, as types.py:60-72."""
    raw = torch.tensor(np.random.standard_normal(size=shape), dtype=TORCH_DTYPE, device=device, requires_grad=True)
    raw.is_trainable = is_trainable
    return PositiveVariable(raw)


def validate_positive(x, dtype=int):
    assert isinstance(dtype, type), 'Specified dtype is not a valid type.'
    assert isinstance(x, dtype), 'The input must be of type {}.'.format(dtype)
    assert (x > 0), 'The input must be positive.'
