import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_params(z):
    from oracle.literal import PARAM_ORDER
    return {k: z["p_" + k] for k in PARAM_ORDER}


MODEL_CASES = ["unit", "t1", "d2t1", "c1", "q10", "mask3", "c3s", "init"]


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def kuu_condition(z):
    """kappa(K_uu + 1e-8 I), worst over the kernel batch.  Two valid float64 evaluation orders of the
    M x M chain differ by ~kappa * eps (SURVEY.md 7 "hard parts"), so parity tolerances state it."""
    import torch
    from oracle import literal as L
    p = {k: torch.as_tensor(v) for k, v in golden_params(z).items()}
    c = L.constrained(p, z["y"].shape[1], int(z["mask_size"]))
    if str(z["mode"]) == "t":
        g, a = c["gamma_atoms"], c["alpha_atoms"]
    else:
        g, a = c["phi"] @ c["gamma_atoms"], c["phi"] @ c["alpha_atoms"]
    k = L.k_uu(c["x_u"], g, a).numpy()
    return max(float(np.linalg.cond(k[i])) for i in range(k.shape[0]))


def tolerances(kappa, base_obj=1e-9, base_grad=1e-9):
    """north_star: 1e-9 relative on the objective and per-block max-norm on gradients, for
    well-conditioned K_uu; widened by the conditioning floor kappa*eps where that is larger."""
    eps = 2.220446049250313e-16
    return max(base_obj, 1e-3 * kappa * eps), max(base_grad, 10.0 * kappa * eps)


# Gradient blocks whose ORACLE value goes through torch's trigamma (autograd of digamma): torch.special.polygamma(1, x)
# carries up to 5e-10 relative error in float64 (its asymptotic series is cut after the x^-7 term; test_oracle_golden.py
# checks this against scipy), and the cancellation in d ELBO / d w_1 amplifies it to ~6e-9.  The product's closed form uses
# a full-precision trigamma (csrc/small.cuh, pinned to 1e-12 by a complex-step derivative in test_gpu_parity.py), so for
# these blocks the comparison with the oracle is held to 5e-8 instead of 1e-9.
TRIGAMMA_BLOCKS = ("gamma1_raw", "gamma2_raw", "w1_raw")


def grad_tol(name, base):
    return max(base, 5e-8) if name in TRIGAMMA_BLOCKS else base
