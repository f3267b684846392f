"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (libdpgp.so)
via the host API that mirrors the reference; the CPU oracle / golden fixtures are only the checker.

Tolerances: north_star asks 1e-9 relative on the objective and per-block max-norm relative on gradients;
`tolerances(kappa)` widens that only by the conditioning floor kappa(K_uu) * eps (c1 fixtures: kappa ~ 1e9)."""
import numpy as np
import pytest
import torch

from conftest import MODEL_CASES, golden_params, grad_tol, kuu_condition, load_golden, report, tolerances

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
    kappa = kuu_condition(z)
    tol_obj, tol_grad = tolerances(kappa)
    ref = float(z["objective"])
    errs = {k: relerr(grads[k], z["g_" + k]) for k in grads if z["g_" + k].size}
    report("%s_%s" % (mode, case), kappa, abs(obj - ref) / abs(ref), errs, tol_obj, tol_grad)
    assert abs(obj - ref) <= tol_obj * abs(ref), (obj, ref)
    for k in errs:
        assert grads[k].shape == z["g_" + k].shape
        assert errs[k] < grad_tol(k, tol_grad), (k, errs[k])
    # accessors
    assert relerr(model.assignments.detach().cpu().numpy(), z["assignments"]) < 1e-14
    assert relerr(model.ard_weights.detach().cpu().numpy(), z["ard_weights"]) < 1e-13
    assert relerr(model.signal_variance.detach().cpu().numpy(), z["signal_variance"]) < 1e-13
    assert relerr(model.noise_precision.detach().cpu().numpy(), z["noise_precision"]) < 1e-13
    assert abs(model.dp.objective.item() - float(z["dp_objective"])) <= 1e-10 * max(1.0, abs(float(z["dp_objective"])))
    xm, xc = model.q_x
    assert tuple(xc.shape) == (z["y"].shape[0], xm.shape[1], xm.shape[1])


def _need_experimental(variant, default_build):
    from dp_gp_lvm_b200 import _lib
    if variant not in default_build and not _lib.has_experimental():
        pytest.skip("variant %d lives in csrc/experimental/ (build with `make EXPERIMENTAL=1`)" % variant)


@pytest.mark.parametrize("exp_variant", [1, 2, 3, 4, 5, 6])
def test_exp_variants_agree(exp_variant):
    _need_experimental(exp_variant, (1, 4))
    z = load_golden("t_q10")
    model = build_model(z, "t", exp_variant=exp_variant)
    obj, grads = model.value_and_grad()
    assert abs(obj - float(z["objective"])) <= 1e-11 * abs(float(z["objective"]))
    for k in grads:
        assert relerr(grads[k], z["g_" + k]) < 1e-9, k


@pytest.mark.parametrize("mode,case", [("t", "q10"), ("d", "c3s"), ("t", "mask3")])
@pytest.mark.parametrize("bwd_variant", [1, 2, 3, 4, 5, 6, 7, 8])
def test_backward_variants_agree(bwd_variant, mode, case):
    """psi2 backward: 1 fused with dD slices, 2 first two-kernel version, 3 fused with tensor-core first phase, 4 fused with
    two 8-warp teams per CTA, 5 fused and warp-specialised (producer / helper warps), 6 fused with dD folded into dZ in the kernel (default),
    7 as 6 with dv / dD as int8 slice products on the tcgen05 tensor cores (csrc/psi2_bwd_umma.cuh), 8 as 6 with the dv / dD contractions as
    FP64 DMMA on the first 8 latent dimensions (csrc/psi2_bwd_mma.cuh) -- all against the reference's gradients.
    The default build holds 6, 1, 7 and 8; the others are built by `make EXPERIMENTAL=1`."""
    _need_experimental(bwd_variant, (1, 6, 7, 8))
    z = load_golden("%s_%s" % (mode, case))
    if bwd_variant == 8 and not 7 <= z["p_x_mean"].shape[1] <= 12:
        pytest.skip("bwd_variant 8 is instantiated for padded Q = 8, 10, 12 only (dpgp_create rejects the others)")
    model = build_model(z, mode, bwd_variant=bwd_variant)
    obj, grads = model.value_and_grad()
    tol_obj, tol_grad = tolerances(kuu_condition(z))
    assert abs(obj - float(z["objective"])) <= tol_obj * abs(float(z["objective"]))
    for k in grads:
        if z["g_" + k].size:
            assert relerr(grads[k], z["g_" + k]) < grad_tol(k, tol_grad), (k, relerr(grads[k], z["g_" + k]))


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
    gp, dstats, dz, dg, da, db, dphi = eng.bound(n, stats, T(zz), gam, alp, bet, phi)
    # the failure also propagates NUMERICALLY (a CUDA-graph replay cannot raise): value and cotangents are NaN
    assert torch.isnan(gp).all() and torch.isnan(dstats[:b * m * m]).any() and torch.isnan(dz).any()
    with pytest.raises(NotPositiveDefiniteError):
        eng.check()
    eng.check()      # flag is cleared after being reported
    # ... and a healthy evaluation afterwards is finite again
    gp2 = eng.bound(n, stats, T(zz + 0.3 * rng.standard_normal(zz.shape)), gam, T(np.ones(b)), bet, phi)[0]
    assert torch.isfinite(gp2).all()
    eng.check()


def test_small_kernels_accept_a_large_truncation_level():
    """dpgp_small_fwd / _bwd with T * (10 + Q) doubles of shared memory beyond the 48 KB default (T = 250, Q = 16)."""
    from dp_gp_lvm_b200.engine import BoundEngine, MODE_T
    rng = np.random.default_rng(3)
    d, t, q = 260, 250, 16
    raw = {"logits": rng.standard_normal((d, t)), "gamma1_raw": rng.standard_normal(t - 1), "gamma2_raw": rng.standard_normal(t - 1),
           "w1_raw": np.array(0.3), "w2_raw": np.array(-0.2), "gamma_atoms_raw": rng.standard_normal((t, q)),
           "alpha_atoms_raw": rng.standard_normal(t), "beta_atoms_raw": rng.standard_normal(t)}
    eng = BoundEngine(300, d, q, 4, t, MODE_T, device=DEV)
    dev = {k: T(v) for k, v in raw.items()}
    phi, gam, alp, bet, scal = eng.small_fwd(dev, t, 1, (1.5, 0.7))
    ref = _np_small_objective(raw, 1.5, 0.7, 1)
    assert abs(float(scal[0] - scal[1]) - ref) <= 1e-12 * abs(ref)
    assert relerr(phi.sum(dim=1).cpu().numpy(), np.ones(d)) < 1e-14


# -------------------------------------------------------------------------------------------- stage level
@pytest.mark.parametrize("shape", [(50, 10, 3, 25, 8), (37, 5, 1, 3, 1), (300, 12, 10, 50, 6), (130, 70, 7, 33, 4),
                                   (96, 64, 10, 128, 2),
                                   # the reference's experiment scripts: Q = 15 (frey_faces_prediction.py:264-266), 16, 20, 22, 25
                                   # (skin_cancer_mnist_tests.py:532-534, :751-753, cmu_swapped_legs_tests.py:22-24), and Q = 13 / 32
                                   (70, 20, 13, 20, 3), (80, 30, 15, 20, 18), (64, 20, 16, 40, 3), (90, 24, 20, 30, 4),
                                   (75, 25, 22, 20, 5), (66, 30, 25, 21, 4), (50, 40, 32, 24, 2),
                                   # M above the tensor-core psi1 path and the 128-column tiles, up to the limit
                                   (70, 9, 10, 129, 2), (40, 6, 4, 160, 2), (50, 8, 10, 200, 2), (48, 6, 6, 256, 1)])
@pytest.mark.parametrize("mode", ["t", "d"])
def test_stages_vs_streaming_oracle(mode, shape):
    """dpgp_stats_fwd / dpgp_bound / dpgp_stats_bwd against oracle/streaming.py on random inputs, incl.
    ragged sizes (N, M not multiples of the tile sizes), Q = 1, T = 1, M = 128, every padded-Q instantiation up to
    Q = 32 and M up to 256."""
    _stages_vs_streaming_oracle(mode, shape, 0)


@pytest.mark.parametrize("shape", [(37, 5, 1, 3, 1), (130, 70, 7, 33, 4), (96, 64, 10, 128, 2), (80, 30, 15, 20, 18), (64, 20, 16, 40, 3),
                                   (200, 12, 10, 50, 6), (1000, 8, 10, 100, 3)])
@pytest.mark.parametrize("mode", ["t", "d"])
def test_stages_vs_streaming_oracle_tensor_core_backward(mode, shape):
    """The same with bwd_variant 7 (csrc/psi2_bwd_umma.cuh: dv and dD as int8 slice products on the tcgen05 tensor cores):
    ragged N (not a multiple of the 64-row item), M off the 8-pair blocks, Q = 1, Q = 15 / 16 (16 accumulator columns), M = 128."""
    _stages_vs_streaming_oracle(mode, shape, 7)


@pytest.mark.parametrize("shape", [(130, 70, 7, 33, 4), (96, 64, 10, 128, 2), (200, 12, 10, 50, 6), (1000, 8, 10, 100, 3), (77, 9, 8, 13, 3),
                                   (90, 14, 12, 40, 2), (150, 6, 11, 24, 5)])
@pytest.mark.parametrize("mode", ["t", "d"])
def test_stages_vs_streaming_oracle_dmma_backward(mode, shape):
    """The same with bwd_variant 8 (csrc/psi2_bwd_mma.cuh: the dv / dD contractions as FP64 DMMA on q < 8, DFMA on the rest):
    padded Q = 8 (no remainder), 10 (two), 12 (four), ragged N and M."""
    _stages_vs_streaming_oracle(mode, shape, 8)


def _stages_vs_streaming_oracle(mode, shape, bwd_variant):
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
    eng = BoundEngine(n, d, q, m, b, MODE_T if mode == "t" else MODE_D, device=DEV, bwd_variant=bwd_variant)
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
    got = {"mu": dmu, "s": ds, "z": dz_k + dz_s, "gamma": dg_k + dg_s, "alpha": da_k + da_s, "beta": dbeta}
    if mode == "t":
        got["phi"] = dphi
    errs = {k_: relerr(v.cpu().numpy().reshape(-1), g_ref[k_].reshape(-1)) for k_, v in got.items()}
    report("stages %s %s bwd_variant %d" % (mode, shape, bwd_variant), kappa, abs(gp.item() - gp_ref) / abs(gp_ref), errs, tol_obj, tol_grad)
    assert abs(gp.item() - gp_ref) <= tol_obj * abs(gp_ref)
    for k_, e in errs.items():
        assert e < tol_grad, (k_, e)


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


@pytest.mark.parametrize("mode,case", [("t", "q10"), ("d", "q10"), ("t", "mask3"), ("d", "mask3"), ("t", "t1"), ("d", "t1"),
                                       ("t", "d2t1"), ("d", "c3s"), ("t", "c1")])
def test_fused_small_kernels_equal_the_torch_chain(mode, case):
    """dpgp_small_fwd / dpgp_small_bwd (softplus / softmax, DP objective with digamma / trigamma closed forms, hyper-prior,
    D-mode mixtures) against the same expressions as torch ops with autograd, and both against the reference's values."""
    z = load_golden("%s_%s" % (mode, case))
    model = build_model(z, mode)
    assert model.fused_small
    obj_f, g_f = model.value_and_grad()
    model.fused_small = False
    obj_t, g_t = model.value_and_grad()
    # the two chains round softplus / softmax differently by an ulp, which an ill-conditioned K_uu (c1: kappa ~ 1e9) amplifies
    kappa = kuu_condition(z)
    tol_obj, tol_grad = tolerances(kappa, base_obj=1e-13, base_grad=1e-11)
    assert abs(obj_f - obj_t) <= tol_obj * abs(obj_t), (obj_f, obj_t)
    for k in g_t:
        if g_t[k].size and np.abs(g_t[k]).max() > 1e-13 * abs(obj_t):      # T = 1: d / d w1_raw vanishes identically (0 vs 2e-16)
            assert relerr(g_f[k], g_t[k]) < grad_tol(k, tol_grad), (k, relerr(g_f[k], g_t[k]))
    assert abs(obj_f - float(z["objective"])) <= max(1e-11, tolerances(kappa)[0]) * abs(float(z["objective"]))


def _np_small_objective(raw, s1, s2, mask):
    """dp objective - hyper-prior of the N-independent variables in numpy / scipy; accepts complex arguments, so that a
    complex-step derivative gives gradients to machine precision, independently of torch's trigamma and of the closed
    forms in csrc/small.cuh.  (dirichlet_process.py:64-88, log_normal.py:24-39, types.py:52-57)"""
    from scipy.special import digamma, loggamma
    sp = lambda x: np.log1p(np.exp(x))
    lg = raw["logits"] - raw["logits"].real.max(axis=1, keepdims=True)
    e = np.exp(lg)
    phi = np.repeat(e / e.sum(axis=1, keepdims=True), mask, axis=0)
    g1, g2, w1, w2 = sp(raw["gamma1_raw"]), sp(raw["gamma2_raw"]), sp(raw["w1_raw"]), sp(raw["w2_raw"])
    t = phi.shape[1]
    d12 = digamma(g1 + g2)
    a, b = digamma(g1) - d12, digamma(g2) - d12
    col = phi.sum(axis=0)
    tail = np.cumsum(col[::-1])[::-1] - col
    c = digamma(w1) - np.log(w2)
    elbo = (col[:-1] * a + tail[:-1] * b).sum() + (t - 1.0) * c + (w1 / w2 - 1.0) * b.sum()
    elbo = elbo + s1 * np.log(s2) - loggamma(s1) + (s1 - 1.0) * c - s2 * w1 / w2
    elbo = elbo - (phi * np.log(phi)).sum()
    elbo = elbo + (loggamma(g1) + loggamma(g2) - loggamma(g1 + g2) - (g1 - 1.0) * digamma(g1) - (g2 - 1.0) * digamma(g2)
                   + (g1 + g2 - 2.0) * d12).sum()
    elbo = elbo + w1 - np.log(w2) + loggamma(w1) + (1.0 - w1) * digamma(w1)
    prior = 0.0
    for k in ("gamma_atoms_raw", "alpha_atoms_raw", "beta_atoms_raw"):
        lx = np.log(sp(raw[k]))
        prior = prior + (-lx - 0.5 * (np.log(2.0 * np.pi) + lx ** 2)).sum()
    return -elbo - prior


@pytest.mark.parametrize("d,t,q,mask", [(6, 5, 3, 1), (12, 4, 2, 3), (7, 2, 1, 1), (40, 20, 10, 1)])
def test_fused_small_gradients_vs_complex_step(d, t, q, mask):
    """dpgp_small_fwd / dpgp_small_bwd with zero bound cotangents: value and every raw gradient of dp - prior against a
    complex-step derivative of the scipy restatement (error ~1e-16, no trigamma involved)."""
    from dp_gp_lvm_b200.engine import BoundEngine, MODE_T
    rng = np.random.default_rng(100 * d + t)
    raw = {"logits": rng.standard_normal((d // mask, t)), "gamma1_raw": rng.standard_normal(t - 1), "gamma2_raw": rng.standard_normal(t - 1),
           "w1_raw": np.array(0.3 + rng.standard_normal()), "w2_raw": np.array(-0.2 + rng.standard_normal()),
           "gamma_atoms_raw": rng.standard_normal((t, q)), "alpha_atoms_raw": rng.standard_normal(t), "beta_atoms_raw": rng.standard_normal(t)}
    s1, s2 = 1.5, 0.7
    eng = BoundEngine(16, d, q, 4, t, MODE_T, device=DEV)
    dev = {k: T(v) for k, v in raw.items()}
    phi, gam, alp, bet, scal = eng.small_fwd(dev, t, mask, (s1, s2))
    ref = _np_small_objective(raw, s1, s2, mask)
    got = float(scal[0] - scal[1])
    assert abs(got - ref) <= 1e-13 * max(1.0, abs(ref)), (got, ref)
    assert relerr(gam.cpu().numpy(), np.log1p(np.exp(raw["gamma_atoms_raw"]))) < 1e-14
    zeros = {"dphi": torch.zeros(d, t, dtype=torch.float64, device=DEV), "dgamma": torch.zeros(t, q, dtype=torch.float64, device=DEV),
             "dalpha": torch.zeros(t, dtype=torch.float64, device=DEV), "dbeta": torch.zeros(t, dtype=torch.float64, device=DEV)}
    names = {"logits": "dlogits", "gamma1_raw": "dgamma1_raw", "gamma2_raw": "dgamma2_raw", "w1_raw": "dw1_raw", "w2_raw": "dw2_raw",
             "gamma_atoms_raw": "dgamma_atoms_raw", "alpha_atoms_raw": "dalpha_atoms_raw", "beta_atoms_raw": "dbeta_atoms_raw"}
    out = {v: torch.zeros_like(dev[k]) for k, v in names.items()}
    eng.small_bwd(dev, t, mask, (s1, s2), phi, zeros["dphi"], zeros["dgamma"], zeros["dalpha"], zeros["dbeta"], out)
    h = 1e-30
    for k, gname in names.items():
        g_ref = np.zeros(raw[k].shape)
        flat = g_ref.reshape(-1)
        for i in range(flat.size):
            pert = {kk: np.array(vv, dtype=np.complex128) for kk, vv in raw.items()}
            pert[k].reshape(-1)[i] += 1j * h
            flat[i] = _np_small_objective(pert, s1, s2, mask).imag / h
        assert relerr(out[gname].cpu().numpy(), g_ref) < 1e-12, (k, relerr(out[gname].cpu().numpy(), g_ref))
