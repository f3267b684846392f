"""TEST INFRASTRUCTURE ONLY -- a TensorFlow-1.x API shim over PyTorch float64 (CPU, eager).

The reference (AndrewRLawrence/dp_gp_lvm) is TensorFlow 1.15 graph code and TensorFlow cannot be
installed in this image (Python 3.12, no network).  This package implements exactly the ~50 `tf.*`
entry points the reference's DP-GP-LVM path touches, with TF-1.15 semantics, so that the UNMODIFIED
reference modules under /root/reference can be imported and executed here:

    sys.path[:0] = ['/root/repo/oracle/tf_shim', '/root/reference']
    from src.models.dp_gp_lvm import dp_gp_lvm_t        # the reference's own code

Graph construction then runs eagerly on torch tensors and `model.objective` is a torch scalar whose
autograd graph follows the reference's op sequence one to one; `oracle/make_golden.py` uses that to
write the fixtures under tests/golden/.  Nothing in the product (`dp_gp_lvm_b200/`) imports this.

`VARIABLE_OVERRIDES` lets the fixture generator evaluate the reference at arbitrary parameter points:
the i-th `tf.Variable` created takes the i-th override instead of its `initial_value` (shapes must
match), because the reference's factories accept no initial values.
"""
import numpy as _np
import torch as _torch

float64 = _torch.float64
float32 = _torch.float32
int32 = _torch.int32
int64 = _torch.int64

VARIABLE_OVERRIDES = None      # list of numpy arrays (or None entries) consumed in creation order
_COLLECTIONS = {"trainable_variables": [], "variables": []}


class GraphKeys:
    TRAINABLE_VARIABLES = "trainable_variables"
    GLOBAL_VARIABLES = "variables"


def reset_default_graph():
    _COLLECTIONS["trainable_variables"].clear()
    _COLLECTIONS["variables"].clear()


def get_collection(key):
    return list(_COLLECTIONS[key])


def _t(x, dtype=None):
    if isinstance(x, _torch.Tensor):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) > 0 and any(isinstance(e, (_torch.Tensor, list, tuple)) for e in x):
        return _torch.stack([_t(e, dtype) for e in x])
    a = _np.asarray(x)
    if dtype is None:
        dtype = _torch.float64 if a.dtype.kind == "f" else None
    return _torch.as_tensor(a, dtype=dtype)


def Variable(initial_value=None, dtype=None, trainable=True, name=None):
    global VARIABLE_OVERRIDES
    idx = len(_COLLECTIONS["variables"])
    if isinstance(initial_value, _torch.Tensor):
        value = initial_value.detach().clone()
    else:
        value = _torch.as_tensor(_np.array(initial_value, dtype=_np.float64))
    if VARIABLE_OVERRIDES is not None and idx < len(VARIABLE_OVERRIDES) and VARIABLE_OVERRIDES[idx] is not None:
        ov = _np.asarray(VARIABLE_OVERRIDES[idx], dtype=_np.float64)
        assert tuple(ov.shape) == tuple(value.shape), (idx, ov.shape, tuple(value.shape))
        value = _torch.as_tensor(ov.copy())
    if dtype is not None:
        value = value.to(dtype)
    value.requires_grad_(True)        # prediction ("non-trainable") variables are optimised too
    _COLLECTIONS["variables"].append(value)
    if trainable:
        _COLLECTIONS["trainable_variables"].append(value)
    return value


def constant(value, dtype=None, shape=None, name=None):
    t = _t(value, dtype)
    if dtype is None and t.dtype == _torch.float32:
        t = t.double()
    return t


def convert_to_tensor(value, dtype=None):
    return _t(value, dtype)


# ----------------------------------------------------------------------------- elementwise
def exp(x): return _torch.exp(_t(x))
def log(x): return _torch.log(_t(x))
def sqrt(x): return _torch.sqrt(_t(x))
def square(x): return _torch.square(_t(x))
def negative(x): return -_t(x)
def reciprocal(x): return 1.0 / _t(x)
def multiply(x, y): return _t(x) * _t(y)
def divide(x, y): return _t(x) / _t(y)
def add(x, y): return _t(x) + _t(y)
def subtract(x, y): return _t(x) - _t(y)
def squared_difference(x, y): return _torch.square(_t(x) - _t(y))
def digamma(x):
    from oracle.special import digamma as _accurate      # torch's own trigamma (autograd of digamma) is only good to ~5e-10
    return _accurate(_t(x))
def lgamma(x): return _torch.lgamma(_t(x))
def abs(x): return _torch.abs(_t(x))
def zeros_like(x, dtype=None): return _torch.zeros_like(_t(x), dtype=dtype)
def ones_like(x, dtype=None): return _torch.ones_like(_t(x), dtype=dtype)
def cast(x, dtype): return _t(x).to(dtype)
def identity(x): return _t(x)
def stop_gradient(x): return _t(x).detach()


def _shape_arg(shape):
    if isinstance(shape, _torch.Tensor):
        return [int(v) for v in shape.reshape(-1).tolist()]
    if isinstance(shape, (int, _np.integer)):
        return [int(shape)]
    return [int(v) for v in shape]


def zeros(shape, dtype=float64): return _torch.zeros(_shape_arg(shape), dtype=dtype)
def ones(shape, dtype=float64): return _torch.ones(_shape_arg(shape), dtype=dtype)


def eye(num_rows, num_columns=None, batch_shape=None, dtype=float64):
    n = int(num_rows)
    m = n if num_columns is None else int(num_columns)
    e = _torch.eye(n, m, dtype=dtype)
    if batch_shape is not None:
        bs = _shape_arg(batch_shape)
        e = e.expand(*bs, n, m)
    return e


def one_hot(indices, depth, dtype=float64):
    idx = _torch.as_tensor(_np.asarray(indices), dtype=_torch.int64)
    return _torch.nn.functional.one_hot(idx, int(depth)).to(dtype)


# ----------------------------------------------------------------------------- shapes
class _Shape(list):
    """tf.shape() result: indexable, each entry a 0-d tensor of the requested dtype."""


def shape(x, out_type=int32):
    return _torch.as_tensor(list(_t(x).shape), dtype=out_type)


def expand_dims(x, axis): return _torch.unsqueeze(_t(x), int(axis))
def squeeze(x, axis=None):
    x = _t(x)
    if axis is None:
        return _torch.squeeze(x)
    axes = [axis] if isinstance(axis, int) else list(axis)
    for a in sorted([a % x.dim() for a in axes], reverse=True):
        x = _torch.squeeze(x, a)
    return x


def reshape(x, shape): return _torch.reshape(_t(x), _shape_arg(shape))
def tile(x, multiples): return _t(x).repeat(*_shape_arg(multiples))


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*perm)


import builtins as _builtins
_pyslice = _builtins.slice


def slice(x, begin, size):
    x = _t(x)
    idx = []
    for b, s in zip(begin, size):
        b = int(b); s = int(s)
        idx.append(_pyslice(b, None if s == -1 else b + s))
    return x[tuple(idx)]


def concat(values, axis): return _torch.cat([_t(v) for v in values], dim=int(axis))
def stack(values, axis=0): return _torch.stack([_t(v) for v in values], dim=int(axis))


# ----------------------------------------------------------------------------- reductions
def _reduce(fn, x, axis, keepdims):
    x = _t(x)
    if axis is None:
        return fn(x)
    return fn(x, dim=axis, keepdim=bool(keepdims))


def reduce_sum(x, axis=None, keepdims=False): return _reduce(_torch.sum, x, axis, keepdims)
def reduce_mean(x, axis=None, keepdims=False): return _reduce(_torch.mean, x, axis, keepdims)
def argmin(x, axis=0): return _torch.argmin(_t(x), dim=int(axis))
def norm(x, axis=None): return _torch.linalg.vector_norm(_t(x), dim=axis)


def cumsum(x, axis=0, exclusive=False, reverse=False):
    x = _t(x)
    axis = int(axis)
    if reverse:
        x = _torch.flip(x, dims=[axis])
    c = _torch.cumsum(x, dim=axis)
    if exclusive:
        c = c - x
    if reverse:
        c = _torch.flip(c, dims=[axis])
    return c


# ----------------------------------------------------------------------------- linear algebra
def matmul(a, b, transpose_a=False, transpose_b=False):
    a = _t(a); b = _t(b)
    if a.dtype != b.dtype:
        b = b.to(a.dtype)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return _torch.matmul(a, b)


def cholesky(x): return _torch.linalg.cholesky(_t(x))


def matrix_triangular_solve(matrix, rhs, lower=True, adjoint=False):
    m = _t(matrix); r = _t(rhs)
    if adjoint:
        m = m.transpose(-1, -2); lower = not lower
    return _torch.linalg.solve_triangular(m, r, upper=not lower)


def matrix_diag(x): return _torch.diag_embed(_t(x))
def matrix_diag_part(x): return _torch.diagonal(_t(x), dim1=-2, dim2=-1)
def diag_part(x): return _torch.diagonal(_t(x), dim1=-2, dim2=-1)
def diag(x): return _torch.diag(_t(x))
def trace(x): return _torch.diagonal(_t(x), dim1=-2, dim2=-1).sum(-1)
def matrix_inverse(x): return _torch.linalg.inv(_t(x))


def map_fn(fn, elems, dtype=None):
    return _torch.stack([fn(e) for e in _t(elems)])


def set_random_seed(seed): _torch.manual_seed(int(seed))


class _NN:
    @staticmethod
    def softmax(x, axis=-1): return _torch.softmax(_t(x), dim=axis)
    @staticmethod
    def softplus(x): return _torch.nn.functional.softplus(_t(x), beta=1.0, threshold=1e9)


nn = _NN()


# ----------------------------------------------------------------------------- session plumbing
class _NoOp:
    pass


def global_variables_initializer(): return _NoOp()
def variables_initializer(var_list=None): return _NoOp()


def _to_numpy(x):
    if isinstance(x, _torch.Tensor):
        return x.detach().cpu().numpy().copy()
    if isinstance(x, (list, tuple)):
        return type(x)(_to_numpy(e) for e in x)
    if isinstance(x, _NoOp) or x is None:
        return None
    return x


class Session:
    def __init__(self, *a, **k): pass
    def run(self, fetches, feed_dict=None): return _to_numpy(fetches)
    def close(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def as_default(self): return self


def get_default_session(): return Session()


# ----------------------------------------------------------------------------- numpy interop
# TF-1 converts numpy operands on either side of an operator.  torch.Tensor declines
# `ndarray (op) Tensor`, so the reflected operators are widened here (this process is test
# infrastructure: the product never imports this module).
def _widen(name):
    orig = getattr(_torch.Tensor, name)

    def op(self, other):
        if isinstance(other, _np.ndarray):
            other = _torch.as_tensor(other, dtype=self.dtype if other.dtype.kind == "f" else None)
        return orig(self, other)
    op.__name__ = name
    setattr(_torch.Tensor, name, op)


for _n in ("__rmul__", "__radd__", "__rsub__", "__rtruediv__", "__mul__", "__add__", "__sub__", "__truediv__"):
    _widen(_n)
