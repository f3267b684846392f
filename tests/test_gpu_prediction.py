"""GPU tests of the prediction paths.  The checker is the reference's own D-mode code
(src/models/dp_gp_lvm.py:234-500) evaluated over the oracle's TF shim at seeded variables: tests/golden/pred_*.npz
(oracle/make_golden.py:prediction_fixture).  T-mode prediction raises NameError upstream; its fixed version is pinned
by the property the reference's own unit test uses for the bound (T-mode == D-mode at equal atoms,
test/unittests/dpgplvm_unitttests.py:547-548)."""
import numpy as np
import pytest
import torch

from conftest import golden_params, kuu_condition, load_golden, tolerances

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def build(z, mode="d", params=None):
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    p = golden_params(z) if params is None else params
    q = p["x_mean"].shape[1]; m = p["x_u"].shape[0]; t = p["gamma_atoms_raw"].shape[0]
    np.random.seed(0)
    if mode == "t":
        model = dp_gp_lvm_t(y_train=z["y"], num_latent_dims=q, num_inducing_points=m, truncation_level=t,
                            alpha_prior_params=z["alpha_prior"], seed=0, device=DEV)
    else:
        model = dp_gp_lvm(y_train=z["y"], num_latent_dims=q, num_inducing_points=m, truncation_level=t,
                          alpha_prior_params=z["alpha_prior"], device=DEV)
    model.load_variables(p)
    return model


@pytest.mark.parametrize("name", ["pred_d_small", "pred_d_q10"])
def test_predict_missing_data_vs_reference(name):
    z = load_golden(name)
    model = build(z)
    do = int(z["d_obs"])
    pred = model.predict_missing_data(y_test=z["y_test"][:, :do])
    pred.load_variables({"x_test_mean": z["xt_mean"], "x_test_var_raw": z["xt_raw"]})
    tol_obj, tol_grad = tolerances(kuu_condition(z))
    lb, xm, xc, mean, covar = pred
    ref = float(z["missing_lower_bound"])
    assert abs(float(lb.item()) - ref) <= tol_obj * abs(ref), (float(lb.item()), ref)
    assert relerr(xc.detach().cpu().numpy(), z["missing_x_test_covar"]) < 1e-14
    assert tuple(mean.shape) == z["missing_predicted_mean"].shape and tuple(covar.shape) == z["missing_predicted_covar"].shape
    assert relerr(mean.cpu().numpy(), z["missing_predicted_mean"]) < 10 * tol_grad
    assert relerr(covar.cpu().numpy(), z["missing_predicted_covar"]) < 10 * tol_grad
    g = torch.autograd.grad(pred.lower_bound, pred.parameters())
    assert relerr(g[0].cpu().numpy(), z["missing_g_xt_mean"]) < tol_grad
    assert relerr(g[1].cpu().numpy(), z["missing_g_xt_raw"]) < tol_grad
    assert float(pred.objective.item()) == -float(pred.lower_bound.item())
    with pytest.raises(AssertionError):
        model.predict_missing_data(y_test=z["y_test"])                       # Do must be < D (dp_gp_lvm.py:322-324)


@pytest.mark.parametrize("name", ["pred_d_small", "pred_d_q10"])
def test_predict_new_latent_variables_vs_reference(name):
    z = load_golden(name)
    model = build(z)
    pred = model.predict_new_latent_variables(y_test=z["y_test"])
    pred.load_variables({"x_test_mean": z["xt_mean"], "x_test_var_raw": z["xt_raw"]})
    tol_obj, tol_grad = tolerances(kuu_condition(z))
    lb, xm, xc, tll = pred
    assert abs(float(lb.item()) - float(z["latent_lower_bound"])) <= tol_obj * abs(float(z["latent_lower_bound"]))
    assert abs(float(tll.item()) - float(z["latent_test_log_likelihood"])) <= tol_obj * abs(float(z["latent_test_log_likelihood"]))
    g = torch.autograd.grad(pred.lower_bound, pred.parameters())
    assert relerr(g[0].cpu().numpy(), z["latent_g_xt_mean"]) < tol_grad
    assert relerr(g[1].cpu().numpy(), z["latent_g_xt_raw"]) < tol_grad
    with pytest.raises(AssertionError):
        model.predict_new_latent_variables(y_test=z["y_test"][:, :3])


def test_t_mode_prediction_equals_d_mode_at_equal_atoms():
    """The reference's T-mode prediction is broken upstream (NameError); the fixed version must reduce to the D-mode
    formulas when all atoms are equal (then every dimension sees the same kernel whatever phi is)."""
    z = load_golden("pred_d_small")
    p = dict(golden_params(z))
    for k in ("gamma_atoms_raw", "alpha_atoms_raw", "beta_atoms_raw"):
        p[k] = np.broadcast_to(p[k][:1], p[k].shape).copy()
    do = int(z["d_obs"])
    outs = []
    for mode in ("d", "t"):
        model = build(z, mode, params=p)
        pred = model.predict_missing_data(y_test=z["y_test"][:, :do], reference_broadcast=False)
        pred.load_variables({"x_test_mean": z["xt_mean"], "x_test_var_raw": z["xt_raw"]})
        lb, _, _, mean, covar = pred
        g = torch.autograd.grad(pred.lower_bound, pred.parameters())
        outs.append((float(lb.item()), mean.cpu().numpy(), covar.cpu().numpy(), g[0].cpu().numpy(), g[1].cpu().numpy()))
    a, b = outs
    assert abs(a[0] - b[0]) <= 1e-10 * abs(a[0])
    for x, y in zip(a[1:], b[1:]):
        assert relerr(y, x) < 1e-9


def test_prediction_optimisation_improves_the_bound_and_the_prediction():
    """test/frey_faces_prediction.py:166-215 in miniature: optimise q(X*) with the second Adam; the missing-data bound
    must rise, and nearest-neighbour initialisation must place test points that ARE training points on their latents."""
    from dp_gp_lvm_b200.train import AdamOptimizer
    z = load_golden("pred_d_q10")
    model = build(z)
    do = int(z["d_obs"])
    np.random.seed(3)
    y_test = z["y"][:6]                                                      # test rows = training rows 0..5
    pred = model.predict_missing_data(y_test=y_test[:, :do])
    x_train = model.q_x[0].detach()
    assert torch.abs(pred.x_test_mean.detach() - x_train[:6]).max().item() < 0.05      # NN index i + N(0, 0.01^2) noise
    lb0 = float(pred.lower_bound.item())
    op = AdamOptimizer(learning_rate=0.02).minimize(loss=pred)
    for _ in range(40):
        op.run()
    pred.engine.check()
    assert float(pred.lower_bound.item()) > lb0
    assert tuple(pred.predicted_mean.shape) == (6, z["y"].shape[1] - do)
