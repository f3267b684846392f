"""Development aid: per-phase device timings of one ELBO+gradient evaluation at a given shape.
python tools/gpu_time.py N D Q M T [mode] [exp_variant] [reps] [bwd_variant]"""
import os
import sys
import json

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dp_gp_lvm_b200.engine import BoundEngine, MODE_D, MODE_T  # noqa: E402

n, d, q, m, t = [int(x) for x in sys.argv[1:6]]
mode = sys.argv[6] if len(sys.argv) > 6 else "t"
expv = int(sys.argv[7]) if len(sys.argv) > 7 else 0
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 3
bwdv = int(sys.argv[9]) if len(sys.argv) > 9 else 0
chv = int(sys.argv[10]) if len(sys.argv) > 10 else 0
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
R = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev, generator=g)
b = t if mode == "t" else d
y = R(n, d); mu = R(n, q); s = torch.exp(0.1 * R(n, q)); z = R(m, q)
gamma = torch.exp(0.3 * R(b, q)); alpha = torch.exp(0.2 * R(b)); beta = 2.0 * torch.exp(0.3 * R(b))
phi = torch.softmax(R(d, t), dim=1).contiguous() if mode == "t" else None
eng = BoundEngine(n, d, q, m, b, MODE_T if mode == "t" else MODE_D, device=dev, exp_variant=expv, bwd_variant=bwdv, chain_variant=chv)
eng.set_timing(True)
print("workspace GB", eng.workspace_bytes / 1e9, flush=True)
for r in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    stats = eng.stats_fwd(mu, s, y, z, gamma, alpha)
    gp, dstats, dz_k, dg_k, da_k, dbeta, dphi = eng.bound(n, stats, z, gamma, alpha, beta, phi)
    dmu, ds, dz_s, dg_s, da_s = eng.stats_bwd(mu, s, y, z, gamma, alpha, dstats)
    e1.record(); torch.cuda.synchronize()
    eng.check()
    tm = eng.timings()
    units = b * n * (m * (m + 1) // 2)
    print(json.dumps({"rep": r, "total_ms": e0.elapsed_time(e1), "gp": float(gp.item()), "phases_ms": tm,
                      "psi2_fwd_tflops_alg71": units * 71 / (tm["psi2_fwd"] * 1e-3) / 1e12,
                      "psi2_fwd_units_per_s": units / (tm["psi2_fwd"] * 1e-3)}), flush=True)
