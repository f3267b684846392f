"""TEST INFRASTRUCTURE ONLY -- makes the UNMODIFIED reference importable in this container.

`activate()` (a) puts `oracle/tf_shim` (TensorFlow-1 API over torch float64) and `/root/reference` on
sys.path, (b) adds a sys.path entry containing 'dp_gp_lvm' because the reference's
`src/utils/constants.py:76` indexes `[p for p in sys.path if 'dp_gp_lvm' in p][-1]`, and (c) restores two
numpy-1.18 behaviours the reference relies on and numpy 2.x removed:
  * `np.int` (used at `src/models/dirichlet_process.py:44`),
  * `np.ones(shape=None)` returning a 0-d array (`src/utils/types.py:52`, called with shape=None at
    `src/models/dirichlet_process.py:58-59`).
Only `oracle/make_golden.py` and `oracle/run_reference_unittests.py` call this, and only here:
`/root/reference` does not exist on the GPU box, so nothing under tests/ may depend on it at run time.
"""
import os
import sys

REFERENCE_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "models"))


def activate():
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
    if not getattr(np.ones, "_dpgp_compat", False):
        _ones = np.ones

        def ones(shape=(), dtype=None, *a, **k):
            return _ones(() if shape is None else shape, dtype, *a, **k)
        ones._dpgp_compat = True
        np.ones = ones
    anchor = os.path.join(_HERE, "_ref", "dp_gp_lvm")      # only its NAME matters (constants.py:76)
    for p in (anchor, REFERENCE_ROOT, os.path.join(_HERE, "tf_shim")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import tensorflow as tf
    assert tf.__file__.startswith(_HERE), "a real tensorflow shadowed the oracle shim"
    return tf
