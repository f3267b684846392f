#!/usr/bin/env python
"""Development aid: bwd_variant 7 (tcgen05 int8 slice products) against bwd_variant 6 (DFMA) on the same inputs:
maximum relative difference of every output of dpgp_stats_bwd, and the device time of the psi2 backward phase.
    python tools/um_check.py N D Q M T [mode] [reps] [variants, e.g. 6,7]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dp_gp_lvm_b200.engine import BoundEngine, MODE_D, MODE_T  # noqa: E402

n, d, q, m, t = [int(x) for x in sys.argv[1:6]]
mode = sys.argv[6] if len(sys.argv) > 6 else "t"
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
variants = [int(v) for v in sys.argv[8].split(",")] if len(sys.argv) > 8 else [6, 7]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(0)
R = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev, generator=g)
b = t if mode == "t" else d
y = R(n, d); mu = R(n, q); s = torch.exp(0.1 * R(n, q)); z = R(m, q)
gamma = torch.exp(0.3 * R(b, q)); alpha = torch.exp(0.2 * R(b)); beta = 2.0 * torch.exp(0.3 * R(b))
phi = torch.softmax(R(d, t), dim=1).contiguous() if mode == "t" else None
out = {}
for variant in variants:
    eng = BoundEngine(n, d, q, m, b, MODE_T if mode == "t" else MODE_D, device=dev, bwd_variant=variant)
    eng.set_timing(True)
    for r in range(reps):
        stats = eng.stats_fwd(mu, s, y, z, gamma, alpha)
        gp, dstats, dz_k, dg_k, da_k, dbeta, dphi = eng.bound(n, stats, z, gamma, alpha, beta, phi)
        res = eng.stats_bwd(mu, s, y, z, gamma, alpha, dstats)
        torch.cuda.synchronize()
        eng.check()
    tm = eng.timings()
    out[variant] = [x.clone() for x in res]
    print("variant %d: psi2 backward %.3f ms, chain %.3f ms (all phases: %s)" % (variant, tm.get("psi2_bwd_fused", float("nan")), tm.get("chain_bwd", float("nan")),
                                                                             {k: round(v, 3) for k, v in tm.items()}), flush=True)
    eng.close()
names = ["dmu", "ds", "dz", "dgamma", "dalpha"]
if len(variants) < 2:
    sys.exit(0)
worst = 0.0
for name, a, c in zip(names, out[variants[0]], out[variants[1]]):
    den = a.abs().max().item()
    err = (a - c).abs().max().item() / max(den, 1e-300)
    worst = max(worst, err)
    print("%-7s max |v7 - v6| / max |v6| = %.3e   (max |v6| %.3e, finite %s)" % (name, err, den, bool(torch.isfinite(c).all())), flush=True)
print("um_check worst %.3e" % worst)
