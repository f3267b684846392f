"""TEST INFRASTRUCTURE -- one-off: re-derives the three gradient blocks of tests/golden/c5_prefix65536.npz that pass
through the trigamma function (gamma1_raw, gamma2_raw, w1_raw) with the accurate derivative of oracle/special.py.

These variables enter the objective only through the DP term (src/models/dirichlet_process.py:64-88), which does not
depend on the data rows, so the 57-minute streaming run of oracle/make_c5_golden.py need not be repeated: the blocks are
d dp_objective / d raw at bench.py's synthetic parameter point.  Every other entry of the fixture is left untouched.

    python -m oracle.patch_c5_trigamma
"""
import os

import numpy as np
import torch

import bench
from oracle import literal as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = os.path.join(ROOT, "tests", "golden", "c5_prefix65536.npz")
    z = dict(np.load(path))
    _, params = bench.synthetic(64, 0, bench.SHAPE)
    names = ("phi_logits", "gamma1_raw", "gamma2_raw", "w1_raw", "w2_raw")
    leaf = {k: torch.tensor(np.asarray(params[k], dtype=np.float64), requires_grad=True) for k in names}
    phi = L.phi_from_logits(leaf["phi_logits"], bench.SHAPE["d"], 1)
    sp = L.softplus
    dp = L.dp_objective(phi, sp(leaf["gamma1_raw"]), sp(leaf["gamma2_raw"]), sp(leaf["w1_raw"]), sp(leaf["w2_raw"]), 1.0, 1.0)
    grads = dict(zip(names, torch.autograd.grad(dp, [leaf[k] for k in names])))
    for k in ("gamma1_raw", "gamma2_raw", "w1_raw"):
        new = grads[k].numpy().reshape(z["grad_" + k].shape)
        rel = np.abs(new - z["grad_" + k]).max() / np.abs(new).max()
        print("%-12s changed by %.2e relative" % (k, rel))
        assert rel < 1e-6
        z["grad_" + k] = new
    # w2_raw also enters only the DP term: it must agree with the stored value to rounding (sanity check of this script)
    assert np.abs(grads["w2_raw"].numpy() - z["grad_w2_raw"]).max() <= 1e-13 * np.abs(z["grad_w2_raw"]).max()
    np.savez_compressed(path, **z)


if __name__ == "__main__":
    main()
