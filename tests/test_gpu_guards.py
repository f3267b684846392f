"""Out-of-bounds checks of our own (compute-sanitizer is closed on the GPU pool this repository is developed on).

  * workspace: with DPGP_GUARD set, every workspace buffer of a handle is bracketed by 4 KB guard bands that
    dpgp_check_guards verifies after the kernels ran (include/dpgp.h);
  * caller-owned buffers: every input and output tensor of the hot-path calls is a slice out of the middle of a larger
    NaN-filled (inputs) / sentinel-filled (outputs) allocation, so a read outside the tensor poisons the result and a write
    outside it disturbs the sentinel.
Shapes are ragged on purpose (N, M, D not multiples of any tile size; Q = 1; M above the shared-memory factor limit)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 777                     # doubles on either side of every caller-owned buffer
SENTINEL = -7.25e300


def padded(a, fill):
    a = np.ascontiguousarray(a, dtype=np.float64)
    buf = torch.full((a.size + 2 * PAD,), fill, dtype=torch.float64, device=DEV)
    view = buf[PAD:PAD + a.size].view(a.shape)
    view.copy_(torch.as_tensor(a, device=DEV))
    return buf, view


def borders_intact(buf, n, fill):
    lo, hi = buf[:PAD], buf[PAD + n:]
    if np.isnan(fill):
        return bool(torch.isnan(lo).all() and torch.isnan(hi).all())
    return bool((lo == fill).all() and (hi == fill).all())


@pytest.mark.parametrize("shape", [(77, 9, 1, 13, 3), (45, 7, 5, 21, 4), (130, 11, 10, 50, 3), (33, 5, 3, 150, 2), (64, 12, 3, 16, 4)])
@pytest.mark.parametrize("mode", ["t", "d"])
def test_no_out_of_bounds_access_on_ragged_shapes(mode, shape, monkeypatch):
    from dp_gp_lvm_b200.engine import MODE_D, MODE_T, BoundEngine
    from oracle import streaming as S
    monkeypatch.setenv("DPGP_GUARD", "1")
    n, d, q, m, t = shape
    b = t if mode == "t" else d
    rng = np.random.default_rng(sum(shape))
    y = rng.standard_normal((n, d)); mu = rng.standard_normal((n, q)); s = np.exp(0.3 * rng.standard_normal((n, q)))
    zz = rng.standard_normal((m, q)) if q > 2 else np.linspace(-3, 3, m)[:, None] * np.ones((1, q)) + 0.05 * rng.standard_normal((m, q))
    gamma = np.exp(0.3 * rng.standard_normal((b, q))); alpha = np.exp(0.2 * rng.standard_normal(b)); beta = 2.0 * np.exp(0.3 * rng.standard_normal(b))
    phi = None
    if mode == "t":
        lg = rng.standard_normal((d, t)); phi = np.exp(lg) / np.exp(lg).sum(1, keepdims=True)
    eng = BoundEngine(n, d, q, m, b, MODE_T if mode == "t" else MODE_D, device=DEV)
    nan = float("nan")
    ins = {k: padded(v, nan) for k, v in dict(mu=mu, s=s, y=y, z=zz, gamma=gamma, alpha=alpha, beta=beta).items()}
    phi_b = padded(phi, nan) if phi is not None else (None, None)
    lib, h = eng.lib, eng._h
    import ctypes as C
    P = lambda t_: C.c_void_p(t_.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mm, mc = m * m, m * (d if mode == "t" else 1)
    outs = {k: padded(np.zeros(sz), SENTINEL) for k, sz in dict(stats=eng.stats_len, gp=1, dstats=eng.stats_len, dz=m * q, dg=b * q, da=b, db=b,
                                                                dphi=d * b, dmu=n * q, ds=n * q, dz2=m * q, dg2=b * q, da2=b).items()}
    o = {k: v[1] for k, v in outs.items()}
    i = {k: v[1] for k, v in ins.items()}
    assert lib.dpgp_stats_fwd(h, P(i["mu"]), P(i["s"]), P(i["y"]), P(i["z"]), P(i["gamma"]), P(i["alpha"]), P(o["stats"]), st) == 0
    assert lib.dpgp_bound(h, n, P(o["stats"]), P(i["z"]), P(i["gamma"]), P(i["alpha"]), P(i["beta"]), P(phi_b[1]) if phi is not None else None,
                          P(o["gp"]), P(o["dstats"]), P(o["dz"]), P(o["dg"]), P(o["da"]), P(o["db"]), P(o["dphi"]) if phi is not None else None, st) == 0
    assert lib.dpgp_stats_bwd(h, P(i["mu"]), P(i["s"]), P(i["y"]), P(i["z"]), P(i["gamma"]), P(i["alpha"]), P(o["dstats"]),
                              P(o["dmu"]), P(o["ds"]), P(o["dz2"]), P(o["dg2"]), P(o["da2"]), st) == 0
    eng.check()
    eng.check_guards()                                                        # workspace guard bands untouched
    for k, (buf, view) in ins.items():
        assert borders_intact(buf, view.numel(), nan), "input %s: border overwritten" % k
    for k, (buf, view) in outs.items():
        assert borders_intact(buf, view.numel(), SENTINEL), "output %s: wrote outside the buffer" % k
    # nothing read the NaN borders of the inputs (or uninitialised workspace): every result is finite and right
    gp_ref, st_ref, g_ref = S.gp_value_and_grad(y, mu, s, zz, gamma, alpha, beta, phi, mode, chunk=32)
    used = [o["stats"], o["gp"], o["dstats"], o["dz"], o["dg"], o["da"], o["db"], o["dmu"], o["ds"], o["dz2"], o["dg2"], o["da2"]]
    assert all(bool(torch.isfinite(t_).all()) for t_ in used)
    assert abs(o["gp"].item() - gp_ref) <= 1e-9 * abs(gp_ref)
    rel = lambda a, r: float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-300))
    # (parity proper lives in test_gpu_parity.py; here a loose, conditioning-aware bound is enough to catch a corrupted result)
    from oracle.literal import k_uu
    kappa = max(np.linalg.cond(k) for k in k_uu(torch.as_tensor(zz), torch.as_tensor(gamma), torch.as_tensor(alpha.reshape(-1, 1))).numpy())
    tol = max(1e-7, 100.0 * kappa * 2.2e-16)
    assert rel(o["dmu"].cpu().numpy().reshape(n, q), g_ref["mu"]) < tol
    assert rel((o["dz"] + o["dz2"]).cpu().numpy().reshape(m, q), g_ref["z"]) < tol
