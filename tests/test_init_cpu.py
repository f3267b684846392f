"""Initialisation of the models (reference src/utils/expressions.py:47-76, src/models/dp_gp_lvm.py:563-575,
src/models/dirichlet_process.py:40-55), on CPU: the factory runs over the oracle-backed stand-in engine (tests/fake_engine.py),
so only host logic is exercised.

The reference's own initial variables are the `p_*` entries of tests/golden/t_init.npz: oracle/make_golden.py built the
UNMODIFIED reference model there with seed 3 and stored what its tf.Variables held (PCA latents, the noisy subset of them
used as inducing inputs, the DP logits and Beta parameters)."""
import numpy as np
import pytest
import torch

from conftest import golden_params, load_golden


def _align_signs(a, ref):
    """Eigenvector signs are arbitrary (ARPACK starts from an unseeded vector) and scipy's `eigs` returns the k leading
    eigenpairs in ARPACK's own order, not sorted by eigenvalue: match every column of `ref` with the column of `a` it is
    (anti-)parallel to, flip signs, and return (aligned a, column permutation)."""
    c = ref.T @ a                                             # [ref column, our column]
    perm = np.abs(c).argmax(axis=1)
    assert len(set(perm.tolist())) == ref.shape[1], "columns do not pair up one to one"
    s = np.sign(c[np.arange(ref.shape[1]), perm])
    return a[:, perm] * s, perm


def test_pca_matches_the_reference_up_to_column_signs():
    """utils/expressions.py obtains the leading eigenvectors of Y Y^T from the D x D matrix Y^T Y; the reference calls ARPACK
    on the N x N matrix (expressions.py:47-76).  Same latents up to sign; ARPACK's tolerance bounds the agreement."""
    from dp_gp_lvm_b200.utils.expressions import principal_component_analysis as pca
    z = load_golden("t_init")
    ref = golden_params(z)["x_mean"]                          # the reference's PCA of z["y"] (N = 60 > D = 12)
    ours, _ = _align_signs(pca(z["y"], ref.shape[1]), ref)
    assert ours.shape == ref.shape
    assert np.abs(ours - ref).max() < 1e-7 * np.abs(ref).max()
    # the other branch (N <= D): against a dense eigen-decomposition of Y Y^T, the quantity the reference asks ARPACK for
    rng = np.random.default_rng(0)
    y = rng.standard_normal((9, 30))
    w, v = np.linalg.eigh(y @ y.T)
    want = v[:, ::-1][:, :4]
    want = want / np.mean(want.std(axis=0, ddof=1))
    got, _ = _align_signs(pca(y, 4), want)
    assert np.abs(got - want).max() < 1e-12
    with pytest.raises(AssertionError):
        pca(y, 9)                                             # Q < min(N, D), the reference's own assertion


@pytest.mark.parametrize("mode", ["t", "d"])
def test_factory_draws_from_numpy_in_the_reference_order(mode):
    """permutation (inducing subset) -> normal (its noise) -> logits -> gamma_1 -> gamma_2 (dp_gp_lvm.py:573-575,
    dirichlet_process.py:40-55): with the reference's seed the initial variables are the reference's, entry by entry."""
    import dp_gp_lvm_b200.models.dp_gp_lvm as M
    import dp_gp_lvm_b200.utils.special as SP
    from fake_engine import OracleEngine, scipy_polygamma
    z = load_golden("t_init")
    ref = golden_params(z)
    n, q = ref["x_mean"].shape
    m, t = ref["x_u"].shape[0], ref["gamma_atoms_raw"].shape[0]
    old = M.ENGINE_FACTORY, SP.POLYGAMMA_HOOK
    M.ENGINE_FACTORY, SP.POLYGAMMA_HOOK = OracleEngine, scipy_polygamma
    try:
        kw = dict(y_train=z["y"], num_latent_dims=q, num_inducing_points=m, truncation_level=t, device="cpu")
        if mode == "t":
            model = M.dp_gp_lvm_t(seed=3, **kw)               # seeds numpy itself, as the reference (:559-560)
        else:
            np.random.seed(3)                                 # D-mode relies on the caller's seed, as the reference
            model = M.dp_gp_lvm(**kw)
    finally:
        M.ENGINE_FACTORY, SP.POLYGAMMA_HOOK = old
    v = {k: t_.detach().numpy() for k, t_ in model.variables.items()}
    x_mean, cols = _align_signs(v["x_mean"], ref["x_mean"])
    assert np.abs(x_mean - ref["x_mean"]).max() < 1e-7 * np.abs(ref["x_mean"]).max()
    # inducing inputs = rows perm[:M] of the PCA latents + N(0, 0.01^2): the permutation and the noise are the reference's
    # (the noise is drawn per (row, latent column) position, so it is compared position by position, not column-matched)
    perm = np.random.RandomState(3).permutation(n)[:m]
    noise_ours = v["x_u"] - v["x_mean"][perm]
    noise_ref = ref["x_u"] - ref["x_mean"][perm]
    assert np.abs(noise_ref).max() < 0.06                     # i.e. the reference did use this permutation
    assert np.abs(noise_ours - noise_ref).max() < 1e-7
    # pure RNG draws: exact
    for k in ("phi_logits", "gamma1_raw", "gamma2_raw"):
        assert v[k].shape == ref[k].shape
        assert np.array_equal(v[k], ref[k]) or np.abs(v[k] - ref[k]).max() < 1e-15, k
    # constants (constants.py:97-99, dp_gp_lvm.py:568-570): q(X) variances and atoms start at 1, (w_1, w_2) at the prior
    for k in ("x_var_raw", "gamma_atoms_raw", "alpha_atoms_raw", "beta_atoms_raw", "w1_raw", "w2_raw"):
        assert np.abs(v[k] - ref[k]).max() < 1e-15, k
