"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (libdpgp.so)
via the host API that mirrors the reference; the CPU oracle / golden fixtures are only the checker.

Tolerances: north_star asks 1e-9 relative on the objective and per-block max-norm relative on gradients;
`tolerances(kappa)` widens that only by the conditioning floor kappa(K_uu) * eps (c1 fixtures: kappa ~ 1e9)."""
import numpy as np
import pytest
import torch

from conftest import MODEL_CASES, golden_params, kuu_condition, load_golden, tolerances

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=DEV)


# ------------------------------------------------------------------------------------------- kernel level
@pytest.mark.parametrize("name", ["kernel_b1", "kernel_b7"])
def test_kernel_statistics_vs_reference(name):
    """k_ard_rbf closures vs the reference's own outputs (test/unittests/kernel_unittests.py, rtol 1e-7 there)."""
    from dp_gp_lvm_b200.kernels.rbf_kernel import k_ard_rbf
    z = load_golden(name)
    k = k_ard_rbf(gamma=z["gamma"], alpha=z["alpha"], beta=z["beta"], device=DEV)
    n, q = z["x_mean"].shape
    x_covar = np.stack([np.diag(z["x_var"][i]) for i in range(n)], axis=0)       # [N,Q,Q] as the reference passes it
    cm = lambda *a, **kw: k.covariance_matrix(*a, **kw).cpu().numpy()
    assert cm(z["x0"], None, include_noise=True, include_jitter=True).shape == z["k_xx"].shape
    assert relerr(cm(z["x0"], None, include_noise=True, include_jitter=True), z["k_xx"]) < 1e-13
    assert relerr(cm(z["x0"], None), z["k_xx_plain"]) < 1e-13
    assert relerr(cm(z["x0"], z["x1"], include_noise=True, include_jitter=True), z["k_xz"]) < 1e-13   # no noise/jitter when input_1 given
    assert relerr(cm(z["x_u"], None, include_jitter=True), z["k_uu"]) < 1e-13
    assert relerr(k.covariance_diag(T(z["x0"]), include_noise=True, include_jitter=True).cpu().numpy(), z["k_diag"]) < 1e-14
    assert relerr(k.psi_0(z["x_u"], T(z["x_mean"]), T(x_covar)).cpu().numpy(), z["psi_0"]) < 1e-15
    p1 = k.psi_1(z["x_u"], z["x_mean"], x_covar).cpu().numpy()
    assert p1.shape == z["psi_1"].shape and relerr(p1, z["psi_1"]) < 1e-12
    p2 = k.psi_2(z["x_u"], z["x_mean"], z["x_var"]).cpu().numpy()                    # [N,Q] diagonal also accepted
    assert p2.shape == z["psi_2"].shape and relerr(p2, z["psi_2"]) < 1e-12
    assert relerr(k.prior_log_likelihood.cpu().numpy(), z["prior_log_likelihood"]) < 1e-13


@pytest.mark.parametrize("name", ["dp_n10_t20", "dp_n12_t5_mask3"])
def test_dirichlet_process_vs_reference(name):
    from dp_gp_lvm_b200.models.dirichlet_process import dirichlet_process
    z = load_golden(name)
    np.random.seed(0)
    dp = dirichlet_process(num_samples=int(z["num_dims"]), alpha_prior_params=z["alpha_prior"],
                           truncation_level=z["phi_logits"].shape[1], mask_size=int(z["mask_size"]), device=DEV)
    leaves = dict(dp.variables)
    with torch.no_grad():
        for k_, v in leaves.items():
            v.copy_(T(z[k_]).reshape(v.shape))
    obj = dp.objective
    assert abs(obj.item() - float(z["objective"])) <= 1e-12 * abs(float(z["objective"]))
    assert relerr(dp.assignments.detach().cpu().numpy(), z["phi"]) < 1e-14
    grads = torch.autograd.grad(obj, list(leaves.values()))
    for k_, g in zip(leaves, grads):
        assert relerr(g.cpu().numpy(), z["grad_" + k_]) < 1e-10, k_


# -------------------------------------------------------------------------------------------- model level
def build_model(z, mode, exp_variant=0, bwd_variant=0):
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    p = golden_params(z)
    n, d = z["y"].shape
    q = p["x_mean"].shape[1]; m = p["x_u"].shape[0]; t = p["gamma_atoms_raw"].shape[0]
    np.random.seed(0)
    if mode == "t":
        model = dp_gp_lvm_t(y_train=z["y"], num_latent_dims=q, num_inducing_points=m, truncation_level=t,
                            alpha_prior_params=z["alpha_prior"], mask_size=int(z["mask_size"]), seed=0, device=DEV,
                            exp_variant=exp_variant, bwd_variant=bwd_variant)
    else:
        model = dp_gp_lvm(y_train=z["y"], num_latent_dims=q, num_inducing_points=m, truncation_level=t,
                          alpha_prior_params=z["alpha_prior"], mask_size=int(z["mask_size"]), device=DEV, exp_variant=exp_variant,
                          bwd_variant=bwd_variant)
    model.load_variables(p)
    return model


@pytest.mark.parametrize("case", MODEL_CASES)
@pytest.mark.parametrize("mode", ["t", "d"])
def test_objective_and_gradients_vs_reference(mode, case):
    """`model.objective` and d objective / d every trainable variable vs the values the reference's own code
    produced for the same variables (tests/golden, oracle/make_golden.py)."""
    z = load_golden("%s_%s" % (mode, case))
    model = build_model(z, mode)
    obj, grads = model.value_and_grad()
    tol_obj, tol_grad = tolerances(kuu_condition(z))
    ref = float(z["objective"])
    assert abs(obj - ref) <= tol_obj * abs(ref), (obj, ref)
    for k in grads:
        if z["g_" + k].size:
            assert grads[k].shape == z["g_" + k].shape
            assert relerr(grads[k], z["g_" + k]) < tol_grad, (k, relerr(grads[k], z["g_" + k]))
    # accessors
    assert relerr(model.assignments.detach().cpu().numpy(), z["assignments"]) < 1e-14
    assert relerr(model.ard_weights.detach().cpu().numpy(), z["ard_weights"]) < 1e-13
    assert relerr(model.signal_variance.detach().cpu().numpy(), z["signal_variance"]) < 1e-13
    assert relerr(model.noise_precision.detach().cpu().numpy(), z["noise_precision"]) < 1e-13
    assert abs(model.dp.objective.item() - float(z["dp_objective"])) <= 1e-10 * max(1.0, abs(float(z["dp_objective"])))
    xm, xc = model.q_x
    assert tuple(xc.shape) == (z["y"].shape[0], xm.shape[1], xm.shape[1])


@pytest.mark.parametrize("exp_variant", [1, 2, 3, 4, 5, 6])
def test_exp_variants_agree(exp_variant):
    z = load_golden("t_q10")
    model = build_model(z, "t", exp_variant=exp_variant)
    obj, grads = model.value_and_grad()
    assert abs(obj - float(z["objective"])) <= 1e-11 * abs(float(z["objective"]))
    for k in grads:
        assert relerr(grads[k], z["g_" + k]) < 1e-9, k


@pytest.mark.parametrize("mode,case", [("t", "q10"), ("d", "c3s"), ("t", "mask3")])
@pytest.mark.parametrize("bwd_variant", [1, 2, 3, 4, 5, 6])
def test_backward_variants_agree(bwd_variant, mode, case):
    """psi2 backward: 1 fused (default), 2 first two-kernel version, 3 fused with tensor-core first phase, 4 fused with
    two 8-warp teams per CTA, 5 fused and warp-specialised (producer / helper warps), 6 fused with dD folded into dZ in the kernel -- all against the reference's gradients."""
    z = load_golden("%s_%s" % (mode, case))
    model = build_model(z, mode, bwd_variant=bwd_variant)
    obj, grads = model.value_and_grad()
    tol_obj, tol_grad = tolerances(kuu_condition(z))
    assert abs(obj - float(z["objective"])) <= tol_obj * abs(float(z["objective"]))
    for k in grads:
        if z["g_" + k].size:
            assert relerr(grads[k], z["g_" + k]) < tol_grad, (k, relerr(grads[k], z["g_" + k]))


def test_t_mode_equals_d_mode_at_equal_atoms():
    """test/unittests/dpgplvm_unitttests.py:547-548 (7 decimals there)."""
    a = build_model(load_golden("t_init"), "t").value_and_grad()[0]
    b = build_model(load_golden("d_init"), "d").value_and_grad()[0]
    assert abs(a - b) < 1e-10 * abs(a)


def test_factory_initialisation_matches_reference_constants():
    """x_var initialised to 1.0, atoms to 1.0 (constants.py:97-99, dp_gp_lvm.py:568-570), inducing inputs a noisy
    subset of the PCA latents (dp_gp_lvm.py:573-575); objective is finite and differentiable."""
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm_t
    rng = np.random.default_rng(5)
    y = rng.standard_normal((80, 9))
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=4, num_inducing_points=20, truncation_level=5, seed=3, device=DEV)
    g, a, b = model.dp_atoms
    assert torch.allclose(g, torch.ones_like(g), atol=1e-14) and torch.allclose(a, torch.ones_like(a), atol=1e-14)
    assert torch.allclose(b, torch.ones_like(b), atol=1e-14)
    xm, xc = model.q_x
    assert torch.allclose(torch.diagonal(xc, dim1=1, dim2=2), torch.ones(80, 4, dtype=torch.float64, device=DEV), atol=1e-14)
    d = torch.cdist(model.inducing_input.detach(), xm.detach()).min(dim=1).values
    assert float(d.max()) < 0.1
    obj = model.objective
    obj.backward()
    assert np.isfinite(obj.item()) and all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_non_positive_definite_is_reported():
    """tf.cholesky fails hard on a non-PD K_uu; here the C ABI reports DPGP_E_NOT_PD with the pivot location."""
    from dp_gp_lvm_b200 import NotPositiveDefiniteError
    from dp_gp_lvm_b200.engine import MODE_T, BoundEngine
    rng = np.random.default_rng(0)
    n, d, q, m, b = 40, 6, 2, 8, 2
    eng = BoundEngine(n, d, q, m, b, MODE_T, device=DEV)
    zz = rng.standard_normal((m, q)); zz[3] = zz[1]                       # duplicate inducing input
    mu, s, y = T(rng.standard_normal((n, q))), T(np.ones((n, q))), T(rng.standard_normal((n, d)))
    gam, alp, bet = T(np.ones((b, q))), T(np.full(b, -1.0)), T(np.ones(b))   # negative signal variance: K_uu not PD
    phi = T(np.full((d, b), 0.5))
    stats = eng.stats_fwd(mu, s, y, T(zz), gam, T(np.ones(b)))
    eng.bound(n, stats, T(zz), gam, alp, bet, phi)
    with pytest.raises(NotPositiveDefiniteError):
        eng.check()
    eng.check()      # flag is cleared after being reported


# -------------------------------------------------------------------------------------------- stage level
@pytest.mark.parametrize("shape", [(50, 10, 3, 25, 8), (37, 5, 1, 3, 1), (300, 12, 10, 50, 6), (130, 70, 7, 33, 4),
                                   (96, 64, 10, 128, 2)])
@pytest.mark.parametrize("mode", ["t", "d"])
def test_stages_vs_streaming_oracle(mode, shape):
    """dpgp_stats_fwd / dpgp_bound / dpgp_stats_bwd against oracle/streaming.py on random inputs, incl.
    ragged sizes (N, M not multiples of the tile sizes), Q = 1, T = 1, M = 128."""
    from dp_gp_lvm_b200.engine import MODE_D, MODE_T, BoundEngine
    from oracle import streaming as S
    n, d, q, m, t = shape
    rng = np.random.default_rng(sum(shape))
    b = t if mode == "t" else d
    y = rng.standard_normal((n, d)); mu = rng.standard_normal((n, q)); s = np.exp(0.3 * rng.standard_normal((n, q)))
    zz = rng.standard_normal((m, q)) * (1.0 if q > 2 else 3.0)
    gamma = np.exp(0.3 * rng.standard_normal((b, q))); alpha = np.exp(0.2 * rng.standard_normal(b)); beta = 2.0 * np.exp(0.3 * rng.standard_normal(b))
    phi = None
    if mode == "t":
        lg = rng.standard_normal((d, t)); phi = np.exp(lg) / np.exp(lg).sum(1, keepdims=True)
    gp_ref, st_ref, g_ref = S.gp_value_and_grad(y, mu, s, zz, gamma, alpha, beta, phi, mode, chunk=32)
    eng = BoundEngine(n, d, q, m, b, MODE_T if mode == "t" else MODE_D, device=DEV)
    args = [T(mu), T(s), T(y), T(zz), T(gamma), T(alpha)]
    stats = eng.stats_fwd(*args)
    gp, dstats, dz_k, dg_k, da_k, dbeta, dphi = eng.bound(n, stats, args[3], args[4], args[5], T(beta), None if phi is None else T(phi))
    dmu, ds, dz_s, dg_s, da_s = eng.stats_bwd(*args, dstats)
    eng.check()
    psi2, pm, yy, kl = eng.split_stats(stats)
    kappa = max(np.linalg.cond(k) for k in __import__("oracle.literal", fromlist=["x"]).k_uu(
        torch.as_tensor(zz), torch.as_tensor(gamma), torch.as_tensor(alpha.reshape(-1, 1))).numpy())
    tol_obj, tol_grad = tolerances(kappa)
    assert relerr(psi2.cpu().numpy(), st_ref["psi2"]) < 1e-12
    assert relerr(pm.cpu().numpy(), st_ref["p"] if mode == "t" else st_ref["p"][:, :, None]) < 1e-12
    assert relerr(yy.cpu().numpy(), st_ref["yy"]) < 1e-13 and relerr(kl.cpu().numpy(), st_ref["kl"]) < 1e-13
    assert abs(gp.item() - gp_ref) <= tol_obj * abs(gp_ref)
    got = {"mu": dmu, "s": ds, "z": dz_k + dz_s, "gamma": dg_k + dg_s, "alpha": da_k + da_s, "beta": dbeta}
    if mode == "t":
        got["phi"] = dphi
    for k_, v in got.items():
        assert relerr(v.cpu().numpy().reshape(-1), g_ref[k_].reshape(-1)) < tol_grad, k_


def test_statistics_are_additive_over_row_shards_and_deterministic():
    """Size-independent properties at a larger N: stats(all rows) == stats(first part) + stats(rest) (the
    identity the N-sharded multi-GPU path relies on), invariance to a row permutation, bitwise run-to-run
    reproducibility, and the value against the chunked CPU oracle."""
    from dp_gp_lvm_b200.engine import MODE_T, BoundEngine
    from oracle import streaming as S
    n, d, q, m, t = 20000, 16, 10, 64, 3
    rng = np.random.default_rng(7)
    y = rng.standard_normal((n, d)); mu = rng.standard_normal((n, q)); s = np.exp(0.2 * rng.standard_normal((n, q)))
    zz = rng.standard_normal((m, q)); gamma = np.exp(0.3 * rng.standard_normal((t, q))); alpha = np.exp(0.2 * rng.standard_normal(t))
    full = BoundEngine(n, d, q, m, t, MODE_T, device=DEV)
    a = full.stats_fwd(T(mu), T(s), T(y), T(zz), T(gamma), T(alpha)).clone()
    b = full.stats_fwd(T(mu), T(s), T(y), T(zz), T(gamma), T(alpha)).clone()
    assert torch.equal(a, b), "not bitwise reproducible"
    k = 7777
    e1 = BoundEngine(k, d, q, m, t, MODE_T, device=DEV); e2 = BoundEngine(n - k, d, q, m, t, MODE_T, device=DEV)
    s1 = e1.stats_fwd(T(mu[:k]), T(s[:k]), T(y[:k]), T(zz), T(gamma), T(alpha))
    s2 = e2.stats_fwd(T(mu[k:]), T(s[k:]), T(y[k:]), T(zz), T(gamma), T(alpha))
    assert relerr((s1 + s2).cpu().numpy(), a.cpu().numpy()) < 1e-12
    perm = rng.permutation(n)
    c = full.stats_fwd(T(mu[perm]), T(s[perm]), T(y[perm]), T(zz), T(gamma), T(alpha))
    assert relerr(c.cpu().numpy(), a.cpu().numpy()) < 1e-12
    with torch.no_grad():
        p2, p = None, None
        for lo in range(0, n, 2000):
            x2, x1 = S.chunk_stats(torch.as_tensor(zz), torch.as_tensor(mu[lo:lo + 2000]), torch.as_tensor(s[lo:lo + 2000]),
                                   torch.as_tensor(y[lo:lo + 2000]), torch.as_tensor(gamma), torch.as_tensor(alpha.reshape(-1, 1)), "t")
            p2 = x2 if p2 is None else p2 + x2; p = x1 if p is None else p + x1
    psi2, pm, yy, kl = full.split_stats(a)
    assert relerr(psi2.cpu().numpy(), p2.numpy()) < 1e-12 and relerr(pm.cpu().numpy(), p.numpy()) < 1e-12
    assert torch.equal(psi2, psi2.transpose(1, 2)), "Psi2 must be exactly symmetric"
