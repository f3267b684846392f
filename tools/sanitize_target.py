#!/usr/bin/env python
"""Target program for compute-sanitizer (profiles/r02_sanitizer.md): one objective + gradient evaluation through the public
API at the smoke shape and at ragged shapes (N, M not multiples of any tile size, Q = 1, M above the shared-memory factor
limit), both modes, default kernel variants, checked against the CPU oracle so that a silent corruption cannot pass.

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_target.py [case ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

CASES = {                     # name: (mode, n, d, q, m, t)
    "smoke": ("t", 64, 12, 3, 16, 4),
    "ragged_t": ("t", 77, 9, 1, 13, 3),
    "ragged_d": ("d", 45, 7, 5, 21, 4),
    "q10_m50": ("t", 100, 20, 10, 50, 5),
    "m150": ("t", 160, 8, 6, 150, 2),
}


def main():
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    from oracle import literal as L
    names = sys.argv[1:] or list(CASES)
    for name in names:
        mode, n, d, q, m, t = CASES[name]
        rng = np.random.default_rng(len(name))
        y = rng.standard_normal((n, d))
        params = L.random_params(rng, n, d, q, m, t)
        if q <= 2:
            params["x_u"] = np.linspace(-2.0, 2.0, m)[:, None] * 3.0 * np.ones((1, q)) + 0.1 * params["x_u"]   # keeps K_uu well conditioned in 1-2 dimensions
        np.random.seed(0)
        if mode == "t":
            model = dp_gp_lvm_t(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, seed=0, device="cuda:0")
        else:
            model = dp_gp_lvm(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, device="cuda:0")
        model.load_variables(params)
        obj, grads = model.value_and_grad()
        ref, gref = L.value_and_grad(L.objective_t if mode == "t" else L.objective_d, y, params)
        err = abs(obj - ref) / abs(ref)
        gerr = max(np.abs(grads[k] - gref[k]).max() / max(np.abs(gref[k]).max(), 1e-300) for k in L.PARAM_ORDER)
        print("%-10s objective rel err %.2e  max gradient rel err %.2e" % (name, err, gerr), flush=True)
        assert err < 1e-8 and gerr < 1e-6, (name, err, gerr)
    torch.cuda.synchronize()
    print("sanitize_target: OK")


if __name__ == "__main__":
    main()
