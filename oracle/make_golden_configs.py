"""TEST INFRASTRUCTURE ONLY -- golden fixtures at the EXACT shapes of the BASELINE.json configurations 2, 3, 4 and of the
headline configuration in D-mode ("What's missing" #1).

  t_c2 / d_c2       (N, D, Q, M, T) = (100, 60, 10, 50, 20)           test/synthetic_data_hard_test.py:21-30,116-118
  t_c3 / d_c3       (300, 60, 10, 50, 10), mask_size 3                test/cmu_walking_tests.py:194-229 (BASELINE shape)
  t_c3m1 / d_c3m1   the same with mask_size 1
      -> the UNMODIFIED reference modules over the TF-1 shim (oracle/make_golden.py: model_fixture), unequal atoms, objective
         and every gradient block; inducing points drawn from the latent means (the reference's initialisation, :573-575).
  c4_t              (1965, 560, 10, 100, 20) T-mode                    test/frey_faces_prediction.py:274-275
  c4_d64            (1965, 64, 10, 100, 20)  D-mode on a 64-column problem of the same N / Q / M / T
  c4_d              (1965, 560, 10, 100, 20) D-mode at the full Frey shape (B = 560 kernels)
  c5_d256           (256, 64, 10, 128, 10)   D-mode, the headline shape on 256 rows
      -> oracle/streaming.py (the [B,N,M,M,Q] tensor of the reference graph does not fit); inputs are regenerated from the
         seed by the test (numpy Generator streams), so only the results are stored.

    python -m oracle.make_golden_configs [name ...]         # runs only in the build container (needs /root/reference)
"""
import os
import sys
import time

import numpy as np
import torch

from oracle.literal import PARAM_ORDER, random_params

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (mode, n, d, q, m, t, mask, seed, z_from_x)
REFERENCE_CASES = {
    "t_c2": ("t", 100, 60, 10, 50, 20, 1, 40, False), "d_c2": ("d", 100, 60, 10, 50, 20, 1, 40, False),
    "t_c3": ("t", 300, 60, 10, 50, 10, 3, 41, True), "d_c3": ("d", 300, 60, 10, 50, 10, 3, 41, True),
    "t_c3m1": ("t", 300, 60, 10, 50, 10, 1, 42, False), "d_c3m1": ("d", 300, 60, 10, 50, 10, 1, 42, False),
}
# name -> (mode, n, d, q, m, t, seed, chunk)
STREAMING_CASES = {
    "c4_t": ("t", 1965, 560, 10, 100, 20, 50, 128),
    "c4_d64": ("d", 1965, 64, 10, 100, 20, 51, 128),
    "c4_d": ("d", 1965, 560, 10, 100, 20, 53, 16),             # all 560 kernels of the D-mode bound: ~25 min on 8 cores
    "c5_d256": ("d", 256, 64, 10, 128, 10, 52, 32),
}


def seeded_problem(seed, n, d, q, m, t):
    """The inputs of a STREAMING_CASES fixture (also called by the tests)."""
    rng = np.random.default_rng(seed)
    y = rng.standard_normal((n, d))
    return y, random_params(rng, n, d, q, m, t)


def streaming_fixture(name):
    from oracle import streaming as S
    mode, n, d, q, m, t, seed, chunk = STREAMING_CASES[name]
    y, params = seeded_problem(seed, n, d, q, m, t)
    t0 = time.perf_counter()
    obj, grads = S.value_and_grad(y, params, mode, chunk=chunk)
    secs = time.perf_counter() - t0
    out = dict(mode=np.array(mode), shape=np.array([n, d, q, m, t]), seed=np.array(seed), objective=np.array(obj),
               seconds=np.array(secs), y_checksum=np.array(float(np.abs(y).sum())))
    out.update({"g_" + k: g for k, g in grads.items()})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote %-10s objective = %.17g  (%.0f s)" % (name, obj, secs), flush=True)


def main():
    names = sys.argv[1:] or list(REFERENCE_CASES) + list(STREAMING_CASES)
    torch.set_num_threads(os.cpu_count() or 1)
    tf = None
    for name in names:
        if name in STREAMING_CASES:
            streaming_fixture(name)
            continue
        if tf is None:
            from oracle import make_golden, ref_env
            tf = ref_env.activate()
        mode, n, d, q, m, t, mask, seed, zfx = REFERENCE_CASES[name]
        t0 = time.perf_counter()
        make_golden.model_fixture(tf, name, mode, n, d, q, m, t, mask_size=mask, seed=seed, z_from_x=zfx)
        print("   (%.0f s)" % (time.perf_counter() - t0), flush=True)


if __name__ == "__main__":
    sys.exit(main())
