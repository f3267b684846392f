#!/usr/bin/env python
"""Development aid: in-situ (warm-cache) device time of every libdpgp kernel of one evaluation, via dpgp_debug_launch_times.
    python tools/launch_times.py small|c4|c5 [rows]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t

which = sys.argv[1] if len(sys.argv) > 1 else "small"
rng = np.random.default_rng(10)
np.random.seed(10)
if which == "small":
    y = rng.standard_normal((100, 60))
    model = dp_gp_lvm(y_train=y, num_latent_dims=10, num_inducing_points=50, truncation_level=20)
elif which == "c3":
    y = rng.standard_normal((300, 60))
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=10, num_inducing_points=50, truncation_level=10, mask_size=3, seed=1)
elif which == "c4":
    y = rng.standard_normal((1965, 560))
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=10, num_inducing_points=100, truncation_level=20, seed=1)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
    y = rng.standard_normal((n, 64))
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=10, num_inducing_points=128, truncation_level=10, seed=1)
params = model.parameters()


def step():
    obj = model.objective
    torch.autograd.grad(obj, params, allow_unused=True)


for _ in range(5):
    step()
eng = model.engine
eng.launch_times(True)
reps = 5
for _ in range(reps):
    step()
rows = eng.launch_times(False)
# one evaluation = the launches from one "(start of dpgp_stats_fwd)" mark to the next
evals = []
for name, us in rows:
    if name.startswith("(start of dpgp_stats_fwd"):
        evals.append([])
    if evals:
        evals[-1].append((name, us))
evals = [e for e in evals if len(e) == len(evals[0])]
s = 0.0
for i, (name, _) in enumerate(evals[0]):
    us = sum(e[i][1] for e in evals) / len(evals)
    print("%3d %-34s %9.1f us" % (i, name, us)); s += us
print("sum %.1f us per evaluation (%d launches, mean of %d evaluations; start marks show the gap before the call)" % (s, len(evals[0]), len(evals)))
