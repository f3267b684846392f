"""Development aid: stage-level parity of the CUDA path against the CPU oracle, printing every error.
Run on a GPU box:  python tools/gpu_check.py [--big]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dp_gp_lvm_b200.engine import BoundEngine, MODE_D, MODE_T  # noqa: E402
from oracle import streaming as S  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def case(n, d, q, m, t, mode, seed=0, exp_variant=0, z_scale=1.0, bwd_variant=0, chain_variant=0):
    rng = np.random.default_rng(seed)
    b = t if mode == "t" else d
    y = rng.standard_normal((n, d))
    mu = rng.standard_normal((n, q)); s = np.exp(0.3 * rng.standard_normal((n, q)))
    z = z_scale * rng.standard_normal((m, q))
    gamma = np.exp(0.3 * rng.standard_normal((b, q))); alpha = np.exp(0.2 * rng.standard_normal(b)); beta = np.exp(0.3 * rng.standard_normal(b)) * 2.0
    phi = None
    if mode == "t":
        lg = rng.standard_normal((d, t)); phi = np.exp(lg) / np.exp(lg).sum(1, keepdims=True)
    t0 = time.time()
    gp_ref, st_ref, g_ref = S.gp_value_and_grad(y, mu, s, z, gamma, alpha, beta, phi, mode, chunk=64)
    t_or = time.time() - t0
    dev = torch.device("cuda:0")
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    eng = BoundEngine(n, d, q, m, b, MODE_T if mode == "t" else MODE_D, device=dev, exp_variant=exp_variant, bwd_variant=bwd_variant, chain_variant=chain_variant)
    mu_d, s_d, y_d, z_d, g_d, a_d, b_d = T(mu), T(s), T(y), T(z), T(gamma), T(alpha), T(beta)
    phi_d = T(phi) if phi is not None else None
    stats = eng.stats_fwd(mu_d, s_d, y_d, z_d, g_d, a_d)
    gp, dstats, dz_k, dg_k, da_k, dbeta, dphi = eng.bound(n, stats, z_d, g_d, a_d, b_d, phi_d)
    dmu, ds, dz_s, dg_s, da_s = eng.stats_bwd(mu_d, s_d, y_d, z_d, g_d, a_d, dstats)
    eng.check()
    torch.cuda.synchronize()
    psi2, pm, yy, kl = eng.split_stats(stats)
    pref = st_ref["p"] if mode == "t" else st_ref["p"][:, :, None]
    out = {
        "psi2": rel(psi2.cpu().numpy(), st_ref["psi2"]), "P": rel(pm.cpu().numpy(), pref), "yy": rel(yy.cpu().numpy(), st_ref["yy"]),
        "kl": rel(kl.cpu().numpy(), st_ref["kl"]), "gp": abs(float(gp.item()) - gp_ref) / abs(gp_ref),
        "dmu": rel(dmu.cpu().numpy(), g_ref["mu"]), "ds": rel(ds.cpu().numpy(), g_ref["s"]),
        "dz": rel((dz_k + dz_s).cpu().numpy(), g_ref["z"]), "dgamma": rel((dg_k + dg_s).cpu().numpy(), g_ref["gamma"]),
        "dalpha": rel((da_k + da_s).cpu().numpy(), g_ref["alpha"].reshape(-1)), "dbeta": rel(dbeta.cpu().numpy(), g_ref["beta"].reshape(-1)),
    }
    if mode == "t":
        out["dphi"] = rel(dphi.cpu().numpy(), g_ref["phi"])
    worst = max(out.values())
    print("[%s] N=%d D=%d Q=%d M=%d B=%d exp=%d bwd=%d ch=%d  worst=%.2e  oracle %.1fs | " % (mode, n, d, q, m, b, exp_variant, bwd_variant, chain_variant, worst, t_or)
          + " ".join("%s=%.1e" % kv for kv in out.items()), flush=True)
    eng.close()
    return worst


if __name__ == "__main__":
    big = "--big" in sys.argv
    worst = 0.0
    worst = max(worst, case(50, 10, 3, 25, 8, "t"))
    worst = max(worst, case(50, 10, 3, 25, 8, "d"))
    worst = max(worst, case(37, 5, 1, 3, 1, "t"))
    worst = max(worst, case(300, 12, 10, 50, 6, "t", seed=1))
    worst = max(worst, case(300, 12, 10, 50, 6, "d", seed=1))
    worst = max(worst, case(100, 20, 5, 25, 8, "t", seed=2, exp_variant=1))
    worst = max(worst, case(100, 20, 5, 25, 8, "t", seed=2, exp_variant=3))
    worst = max(worst, case(130, 70, 7, 33, 4, "t", seed=3))
    worst = max(worst, case(130, 70, 7, 33, 4, "t", seed=3, bwd_variant=2, chain_variant=2))
    worst = max(worst, case(130, 70, 7, 33, 4, "t", seed=3, bwd_variant=4))
    worst = max(worst, case(300, 12, 10, 50, 6, "d", seed=1, bwd_variant=4))
    worst = max(worst, case(37, 5, 1, 3, 1, "t", bwd_variant=4))
    worst = max(worst, case(77, 12, 16, 40, 3, "t", seed=6, bwd_variant=4))
    worst = max(worst, case(130, 70, 7, 33, 4, "t", seed=3, bwd_variant=3))
    worst = max(worst, case(300, 12, 10, 50, 6, "d", seed=1, bwd_variant=3))
    worst = max(worst, case(100, 20, 5, 25, 8, "t", seed=2, bwd_variant=3, exp_variant=2))
    worst = max(worst, case(37, 5, 1, 3, 1, "t", bwd_variant=3))
    worst = max(worst, case(65, 9, 8, 140, 2, "t", seed=7, bwd_variant=3))
    worst = max(worst, case(130, 70, 7, 33, 4, "t", seed=3, bwd_variant=1))
    worst = max(worst, case(300, 12, 10, 50, 6, "d", seed=1, bwd_variant=1))
    worst = max(worst, case(37, 5, 1, 3, 1, "t", bwd_variant=1))
    worst = max(worst, case(65, 9, 8, 140, 2, "t", seed=7, bwd_variant=1))
    worst = max(worst, case(77, 12, 16, 40, 3, "t", seed=6, bwd_variant=1))
    worst = max(worst, case(130, 70, 7, 33, 4, "t", seed=3, bwd_variant=5))
    worst = max(worst, case(300, 12, 10, 50, 6, "d", seed=1, bwd_variant=5))
    worst = max(worst, case(37, 5, 1, 3, 1, "t", bwd_variant=5))
    worst = max(worst, case(65, 9, 8, 140, 2, "t", seed=7, bwd_variant=5))
    worst = max(worst, case(100, 20, 5, 25, 8, "t", seed=2, bwd_variant=5, exp_variant=2))
    worst = max(worst, case(130, 70, 7, 33, 4, "t", seed=3, z_scale=30.0))
    worst = max(worst, case(90, 150, 4, 20, 5, "t", seed=8))
    worst = max(worst, case(100, 20, 5, 25, 8, "t", seed=2, exp_variant=2))
    worst = max(worst, case(300, 12, 10, 50, 6, "d", seed=1, exp_variant=2, bwd_variant=2))
    worst = max(worst, case(100, 20, 5, 25, 8, "t", seed=2, exp_variant=5))
    worst = max(worst, case(300, 12, 10, 50, 6, "d", seed=1, exp_variant=6))
    worst = max(worst, case(77, 12, 16, 40, 3, "t", seed=6))
    worst = max(worst, case(65, 9, 8, 140, 2, "t", seed=7))
    if big:
        worst = max(worst, case(200, 64, 10, 128, 3, "t", seed=4))
        worst = max(worst, case(200, 64, 10, 128, 3, "t", seed=4, bwd_variant=1))
        worst = max(worst, case(200, 64, 10, 128, 3, "t", seed=4, bwd_variant=3))
        worst = max(worst, case(200, 64, 10, 128, 3, "t", seed=4, bwd_variant=4))
        worst = max(worst, case(200, 64, 10, 128, 3, "t", seed=4, bwd_variant=5))
        worst = max(worst, case(1000, 16, 10, 100, 2, "d", seed=5))
    print("WORST", worst)
    sys.exit(0 if worst < 1e-8 else 1)
