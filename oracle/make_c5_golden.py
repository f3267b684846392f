#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- generates tests/golden/c5_prefix<N>.npz (headline shape C5 on a row prefix).

Runs the CPU streaming oracle (oracle/streaming.py, the literal float64 restatement of dp_gp_lvm.py:582-676 evaluated
chunk by chunk) on the first N rows of bench.py's synthetic headline problem (D = 64, Q = 10, M = 128, T = 10, seed 0)
and stores the objective, every small gradient block in full, and the per-row gradient blocks (x_mean, x_var_raw) as
sums, norms and a strided row sample.  N = 65 536 takes ~45 min on 8 cores; the fixture is committed so the GPU box
(which has no /root/reference and no hours to spare) only compares.

    python oracle/make_c5_golden.py [N] [threads]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STRIDE = 257                       # rows kept from the per-row gradient blocks


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    import bench
    from oracle import streaming as S
    torch.set_num_threads(threads)
    y, params = bench.synthetic(n, 0, bench.SHAPE)
    t0 = time.perf_counter()
    obj, grads = S.value_and_grad(y, params, "t", chunk=128)
    secs = time.perf_counter() - t0
    out = {"n": np.array(n), "objective": np.array(obj), "seconds": np.array(secs), "stride": np.array(STRIDE)}
    for k, g in grads.items():
        if k in ("x_mean", "x_var_raw"):
            out["sum_" + k] = g.sum(axis=0)
            out["norm_" + k] = np.array(np.sqrt((g ** 2).sum()))
            out["rows_" + k] = g[::STRIDE].copy()
        else:
            out["grad_" + k] = g
    path = os.path.join(ROOT, "tests", "golden", "c5_prefix%d.npz" % n)
    np.savez_compressed(path, **out)
    print("wrote %s: objective %.17g in %.0f s" % (path, obj, secs))


if __name__ == "__main__":
    main()
