"""Development aid: a few eager Adam iterations at the reference's small shapes (for an ncu launch list)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
from dp_gp_lvm_b200.train import AdamOptimizer

mode = sys.argv[1] if len(sys.argv) > 1 else "d"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(10)
y = rng.standard_normal((100, 60))
np.random.seed(10)
if mode == "d":
    model = dp_gp_lvm(y_train=y, num_latent_dims=10, num_inducing_points=50, truncation_level=20)
else:
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=10, num_inducing_points=50, truncation_level=20, seed=10)
op = AdamOptimizer(learning_rate=0.01).minimize(loss=model)
for _ in range(iters):
    op.run()
torch.cuda.synchronize()
print("objective", float(op.objective.item()))
