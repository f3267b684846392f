#!/usr/bin/env python
"""Development aid: CUDA-graph capture of one Adam iteration at BASELINE.json configs[0..3], both modes, with and without
eager evaluations of another model of the same shape beforehand (bench.py's configs block does the latter)."""
import gc
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
from dp_gp_lvm_b200.train import AdamOptimizer

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
only = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]
for idx, (name, src, n, d, q, m, t, mask) in enumerate(bench.CONFIGS):
    if idx not in only:
        continue
    shape = dict(n=n, d=d, q=q, m=m, t=t, mask=mask)
    y, params = bench.synthetic(n, 0, shape, seed=100 + idx)
    for mode in ("t", "d"):
        for eager_first in (False, True):
            kw = dict(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, mask_size=mask, device=dev)

            def build():
                np.random.seed(0)
                mdl = dp_gp_lvm_t(seed=0, **kw) if mode == "t" else dp_gp_lvm(**kw)
                mdl.load_variables(params)
                return mdl
            try:
                if eager_first:
                    model = build()
                    leaves = model.parameters()
                    for _ in range(3):
                        obj = model.objective
                        torch.autograd.grad(obj, leaves, allow_unused=True)
                    model.engine.check()
                    del model, leaves, obj
                model = build()
                op = AdamOptimizer(learning_rate=0.01, use_cuda_graph=True).minimize(loss=model)
                for _ in range(3):
                    op.run()
                torch.cuda.synchronize()
                model.engine.check()
                print("OK   %s mode %s eager_first %s objective %.6f" % (name, mode, eager_first, float(op.objective.item())), flush=True)
            except Exception as e:
                print("FAIL %s mode %s eager_first %s: %s" % (name, mode, eager_first, str(e).splitlines()[0]), flush=True)
                traceback.print_exc()
                try:
                    print("last lib error:", model.engine.lib.dpgp_last_error(model.engine._h).decode(), flush=True)
                except Exception:
                    pass
                try:
                    torch.cuda.synchronize()
                except Exception as e2:
                    print("sync after failure:", str(e2).splitlines()[0], flush=True)
            model = op = None
            gc.collect()
            torch.cuda.empty_cache()
