#!/usr/bin/env python
"""Development aid: bisects the capture failure seen in bench.py (a graph-mode TrainOp of one model, then the capture of another)."""
import gc
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VARIANTS = ["eager_only", "graph_only", "both", "both_gc", "graph_keepalive", "graph_only_noinline_check"]


def child(variant):
    import numpy as np
    import torch
    import bench
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    from dp_gp_lvm_b200.train import AdamOptimizer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    rng = np.random.default_rng(10)
    y = rng.standard_normal((100, 60))
    keep = []

    def first(graph):
        np.random.seed(10)
        model = dp_gp_lvm(y_train=y, num_latent_dims=10, num_inducing_points=50, truncation_level=20, device=dev)
        op = AdamOptimizer(learning_rate=0.01, use_cuda_graph=graph).minimize(loss=model)
        for _ in range(8):
            op.run()
        torch.cuda.synchronize()
        float(op.objective.item())
        if variant != "graph_only_noinline_check":
            model.engine.check()
        if variant == "graph_keepalive":
            keep.append((model, op))

    if variant in ("eager_only", "both", "both_gc"):
        first(False)
    if variant in ("graph_only", "both", "both_gc", "graph_keepalive", "graph_only_noinline_check"):
        first(True)
    if variant == "both_gc":
        gc.collect(); torch.cuda.empty_cache()
    name, src, n, d, q, m, t, mask = bench.CONFIGS[0]
    shape = dict(n=n, d=d, q=q, m=m, t=t, mask=mask)
    y0, params = bench.synthetic(n, 0, shape, seed=100)
    np.random.seed(0)
    model = dp_gp_lvm_t(seed=0, y_train=y0, num_latent_dims=q, num_inducing_points=m, truncation_level=t, mask_size=mask, device=dev)
    model.load_variables(params)
    op = AdamOptimizer(learning_rate=0.01, use_cuda_graph=True).minimize(loss=model)
    for _ in range(3):
        op.run()
    torch.cuda.synchronize()
    print("objective %.6f" % float(op.objective.item()))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for v in VARIANTS:
            r = subprocess.run([sys.executable, __file__, v], capture_output=True, text=True)
            err = [l for l in r.stderr.splitlines() if "Error" in l or "error" in l]
            print("%-28s rc %d  %s  %s" % (v, r.returncode, r.stdout.strip(), err[-1] if err and r.returncode else ""), flush=True)
