"""GPU tests of the Bayesian GP-LVM: the B = 1 case of the streamed bound against the reference's own
`bayesian_gp_lvm` (src/models/gaussian_process.py:132-548) evaluated over the oracle's TF shim
(tests/golden/bgplvm_q4.npz, oracle/make_golden.py:bgplvm_fixture)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ("gamma_raw", "alpha_raw", "beta_raw", "x_mean", "x_u", "x_var_raw")


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def build(z):
    from dp_gp_lvm_b200.models.gaussian_process import bayesian_gp_lvm
    np.random.seed(0)
    model = bayesian_gp_lvm(y_train=z["y"], num_latent_dims=z["p_x_mean"].shape[1], num_inducing_points=z["p_x_u"].shape[0], device=DEV)
    model.load_variables({k: z["p_" + k] for k in NAMES})
    return model


def test_objective_and_gradients_vs_reference():
    z = load_golden("bgplvm_q4")
    model = build(z)
    obj, grads = model.value_and_grad()
    assert abs(obj - float(z["objective"])) <= 1e-9 * abs(float(z["objective"]))
    for k in NAMES:
        assert relerr(grads[k], z["g_" + k]) < 1e-9, (k, relerr(grads[k], z["g_" + k]))
    assert relerr(model.ard_weights.detach().cpu().numpy(), z["ard_weights"]) < 1e-14
    assert relerr(model.noise_precision.detach().cpu().numpy(), z["noise_precision"]) < 1e-14
    xm, xc = model.q_x
    assert tuple(xc.shape) == (z["y"].shape[0], xm.shape[1], xm.shape[1])


def test_factory_assertions_and_initialisation():
    from dp_gp_lvm_b200.models.gaussian_process import bayesian_gp_lvm
    y = np.random.default_rng(0).standard_normal((30, 6))
    with pytest.raises(AssertionError):
        bayesian_gp_lvm(y_train=y, num_latent_dims=6, num_inducing_points=10, device=DEV)       # Q < D
    with pytest.raises(AssertionError):
        bayesian_gp_lvm(y_train=y, num_latent_dims=2, num_inducing_points=30, device=DEV)       # M < N (strict, :156)
    with pytest.raises(NotImplementedError):
        bayesian_gp_lvm(y_train=y, num_latent_dims=2, num_inducing_points=10, num_latent_samples=5, device=DEV)
    np.random.seed(1)
    model = bayesian_gp_lvm(y_train=y, num_latent_dims=2, num_inducing_points=10, device=DEV)
    xm, xc = model.q_x
    assert torch.allclose(torch.diagonal(xc, dim1=1, dim2=2), torch.full((30, 2), 0.5, dtype=torch.float64, device=DEV), atol=1e-14)
    assert np.isfinite(float(model.objective.item()))


def test_predictions_vs_reference():
    z = load_golden("bgplvm_q4")
    model = build(z)
    do = int(z["d_obs"])
    pred = model.predict_missing_data(y_test=z["y_test"][:, :do])
    pred.load_variables({"x_test_mean": z["xt_mean"], "x_test_var_raw": z["xt_raw"]})
    lb, xm, xc, mean, covar = pred
    assert abs(float(lb.item()) - float(z["missing_lower_bound"])) <= 1e-9 * abs(float(z["missing_lower_bound"]))
    assert relerr(mean.cpu().numpy(), z["missing_predicted_mean"]) < 1e-8
    assert relerr(covar.cpu().numpy(), z["missing_predicted_covar"]) < 1e-8
    g = torch.autograd.grad(pred.lower_bound, pred.parameters())
    assert relerr(g[0].cpu().numpy(), z["missing_g_xt_mean"]) < 1e-9 and relerr(g[1].cpu().numpy(), z["missing_g_xt_raw"]) < 1e-9
    pred2 = model.predict_new_latent_variables(y_test=z["y_test"])
    pred2.load_variables({"x_test_mean": z["xt_mean"], "x_test_var_raw": z["xt_raw"]})
    lb2, _, _, tll = pred2
    assert abs(float(lb2.item()) - float(z["latent_lower_bound"])) <= 1e-9 * abs(float(z["latent_lower_bound"]))
    assert abs(float(tll.item()) - float(z["latent_test_log_likelihood"])) <= 1e-9 * abs(float(z["latent_test_log_likelihood"]))


def test_bgplvm_trains_and_saves(tmp_path):
    from dp_gp_lvm_b200.train import save_results, train
    z = load_golden("bgplvm_q4")
    model = build(z)
    t_opt, hist = train(model, learning_rate=0.01, train_iter=11, print_every=5, verbose=False, name='BGP-LVM')
    assert hist[-1][1] < hist[0][1]
    out = save_results(model, str(tmp_path / "bgplvm.npz"), z["y"], train_opt_time=t_opt)
    assert "assignments" not in out and out["ard_weights"].shape == (1, z["p_x_mean"].shape[1])
