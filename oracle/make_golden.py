"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from THE REFERENCE ITSELF.

Runs only in the build container (needs /root/reference).  The reference's unmodified modules
(`src/kernels/rbf_kernel.py`, `src/models/dirichlet_process.py`, `src/models/dp_gp_lvm.py`) are imported
over the oracle's TensorFlow-1 shim (oracle/tf_shim, validated by oracle/run_reference_unittests.py) and
evaluated at seeded parameter points; objective values come from the reference's own op sequence and
gradients from torch autograd through that same sequence.  The fixtures are small and committed, so the
tests on the GPU box never need /root/reference.

    python -m oracle.make_golden          # regenerates every fixture (seeded; only t_init/d_init depend on
                                          # ARPACK's unseeded PCA start vector and store their own inputs)
"""
import os
import sys

import numpy as np
import torch

from oracle import ref_env
from oracle.literal import PARAM_ORDER, random_params

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(x):
    return x.detach().numpy().copy() if isinstance(x, torch.Tensor) else np.asarray(x)


def kernel_fixture(tf, name, batch, seed=1, n=100, n1=75, m=25, q=10):
    """Inputs drawn as test/unittests/kernel_unittests.py:158-179 (B=1) and :477-498 (B=7)."""
    from src.kernels.rbf_kernel import k_ard_rbf
    rs = np.random.RandomState(seed)
    x0 = rs.standard_normal((n, q)); x1 = rs.standard_normal((n1, q))
    x_mean = rs.standard_normal((n, q)); x_var = np.square(rs.standard_normal((n, q)))
    x_covar = np.stack([np.diag(x_var[i]) for i in range(n)], axis=0)
    x_u = rs.standard_normal((m, q))
    gamma = np.exp(rs.standard_normal((batch, q)))
    alpha = np.square(rs.standard_normal((batch, 1)) + 1.0)
    beta = np.square(rs.standard_normal((batch, 1)) + np.sqrt(50.0))
    tf.reset_default_graph()
    k = k_ard_rbf(gamma=tf.constant(gamma), alpha=tf.constant(alpha), beta=tf.constant(beta))
    out = dict(
        x0=x0, x1=x1, x_mean=x_mean, x_var=x_var, x_u=x_u, gamma=gamma, alpha=alpha, beta=beta,
        k_xx=_np(k.covariance_matrix(x0, None, include_noise=True, include_jitter=True)),
        k_xx_plain=_np(k.covariance_matrix(x0, None, include_noise=False, include_jitter=False)),
        k_xz=_np(k.covariance_matrix(x0, x1, include_noise=True, include_jitter=True)),
        k_uu=_np(k.covariance_matrix(x_u, None, include_noise=False, include_jitter=True)),
        k_diag=_np(k.covariance_diag(x0, include_noise=True, include_jitter=True)),
        psi_0=_np(k.psi_0(x_u, x_mean, x_covar)), psi_1=_np(k.psi_1(x_u, x_mean, x_covar)),
        psi_2=_np(k.psi_2(x_u, x_mean, x_covar)), prior_log_likelihood=_np(k.prior_log_likelihood))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "psi_2[0,0,:2] =", out["psi_2"][0, 0, :2])


def dp_fixture(tf, name, d, t, mask_size, seed, alpha_prior):
    from src.models.dirichlet_process import dirichlet_process
    rng = np.random.default_rng(seed)
    ov = [rng.standard_normal((d // mask_size, t)), rng.standard_normal(t - 1), rng.standard_normal(t - 1),
          np.array(rng.standard_normal()), np.array(rng.standard_normal())]
    tf.reset_default_graph()
    tf.VARIABLE_OVERRIDES = ov
    dp = dirichlet_process(num_samples=d, alpha_prior_params=np.array(alpha_prior), truncation_level=t,
                           mask_size=mask_size)
    tf.VARIABLE_OVERRIDES = None
    vs = tf.get_collection("variables")
    grads = torch.autograd.grad(dp.objective, vs)
    names = ("phi_logits", "gamma1_raw", "gamma2_raw", "w1_raw", "w2_raw")
    out = {n_: v for n_, v in zip(names, ov)}
    out.update({"grad_" + n_: _np(g) for n_, g in zip(names, grads)})
    out.update(objective=_np(dp.objective), phi=_np(dp.q_z), alpha_prior=np.array(alpha_prior),
               mask_size=np.array(mask_size), num_dims=np.array(d))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "objective =", float(dp.objective))


def model_fixture(tf, name, mode, n, d, q, m, t, mask_size=1, seed=0, z_from_x=False, alpha_prior=(1.0, 1.0),
                  override=True, params=None):
    from src.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    rng = np.random.default_rng(seed)
    y = rng.standard_normal((n, d))
    if params is None:
        params = random_params(rng, n, d, q, m, t, mask_size=mask_size, z_from_x=z_from_x)
    tf.reset_default_graph()
    tf.VARIABLE_OVERRIDES = [params[k] for k in PARAM_ORDER] if override else None
    np.random.seed(seed)
    if mode == "t":
        model = dp_gp_lvm_t(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t,
                            alpha_prior_params=np.array(alpha_prior), mask_size=mask_size, seed=seed)
    else:
        model = dp_gp_lvm(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t,
                          alpha_prior_params=np.array(alpha_prior), mask_size=mask_size)
    tf.VARIABLE_OVERRIDES = None
    vs = tf.get_collection("variables")
    assert len(vs) == len(PARAM_ORDER)
    if not override:
        params = {k: _np(v) for k, v in zip(PARAM_ORDER, vs)}
    grads = torch.autograd.grad(model.objective, vs)
    out = dict(y=y, mode=np.array(mode), mask_size=np.array(mask_size), alpha_prior=np.array(alpha_prior),
               objective=_np(model.objective), dp_objective=_np(model.dp.objective),
               assignments=_np(model.assignments), ard_weights=_np(model.ard_weights),
               signal_variance=_np(model.signal_variance), noise_precision=_np(model.noise_precision))
    for k, g in zip(PARAM_ORDER, grads):
        out["p_" + k] = np.asarray(params[k], dtype=np.float64)
        out["g_" + k] = _np(g)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote %-14s objective = %.15g" % (name, float(model.objective)))
    return float(model.objective), params


def prediction_fixture(tf, name, n, d, q, m, t, n_test, d_obs, seed):
    """The reference's D-mode prediction graphs (src/models/dp_gp_lvm.py:234-500) at seeded training variables and
    seeded q(X*) variables: values, predictive moments, and the gradients of the two lower bounds w.r.t. q(X*)."""
    from src.models.dp_gp_lvm import dp_gp_lvm
    rng = np.random.default_rng(seed)
    y = rng.standard_normal((n, d))
    params = random_params(rng, n, d, q, m, t)
    y_test = rng.standard_normal((n_test, d))
    xt_mean = rng.standard_normal((n_test, q)); xt_raw = 0.4 + 0.2 * rng.standard_normal((n_test, q))
    out = dict(y=y, mode=np.array("d"), mask_size=np.array(1), alpha_prior=np.array([1.0, 1.0]), y_test=y_test,
               d_obs=np.array(d_obs), xt_mean=xt_mean, xt_raw=xt_raw)
    for k in PARAM_ORDER:
        out["p_" + k] = np.asarray(params[k], dtype=np.float64)
    for which in ("missing", "latent"):
        tf.reset_default_graph()
        tf.VARIABLE_OVERRIDES = [params[k] for k in PARAM_ORDER]
        np.random.seed(seed)
        model = dp_gp_lvm(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t,
                          alpha_prior_params=np.array([1.0, 1.0]))
        tf.VARIABLE_OVERRIDES = [None] * len(PARAM_ORDER) + [xt_mean, xt_raw]
        if which == "missing":
            lb, xm, xc, pm, pc = model.predict_missing_data(y_test=y_test[:, :d_obs])
            out.update(missing_predicted_mean=_np(pm), missing_predicted_covar=_np(pc))
        else:
            lb, xm, xc, tll = model.predict_new_latent_variables(y_test=y_test)
            out.update(latent_test_log_likelihood=_np(tll))
        tf.VARIABLE_OVERRIDES = None
        vs = tf.get_collection("variables")
        assert len(vs) == len(PARAM_ORDER) + 2
        g = torch.autograd.grad(lb, vs[-2:])
        out.update({which + "_lower_bound": _np(lb), which + "_g_xt_mean": _np(g[0]), which + "_g_xt_raw": _np(g[1]),
                    which + "_x_test_covar": _np(xc)})
    out["objective"] = _np(model.objective)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote %-14s missing lb = %.15g, latent lb = %.15g" % (name, float(out["missing_lower_bound"]), float(out["latent_lower_bound"])))


def bgplvm_fixture(tf, name, n, d, q, m, n_test, d_obs, seed):
    """The reference's Bayesian GP-LVM (src/models/gaussian_process.py:132-548): objective + gradients at seeded
    variables, and its two prediction graphs at seeded q(X*)."""
    from src.models.gaussian_process import bayesian_gp_lvm
    rng = np.random.default_rng(seed)
    y = rng.standard_normal((n, d)); y_test = rng.standard_normal((n_test, d))
    base = float(np.log(np.expm1(1.0)))
    names = ("gamma_raw", "alpha_raw", "beta_raw", "x_mean", "x_u", "x_var_raw")
    params = dict(gamma_raw=base + 0.3 * rng.standard_normal((1, q)), alpha_raw=base + 0.3 * rng.standard_normal((1, 1)),
                  beta_raw=base + 0.3 * rng.standard_normal((1, 1)), x_mean=rng.standard_normal((n, q)),
                  x_u=rng.standard_normal((m, q)), x_var_raw=0.3 * rng.standard_normal((n, q)))
    xt_mean = rng.standard_normal((n_test, q)); xt_raw = 0.4 + 0.2 * rng.standard_normal((n_test, q))
    out = dict(y=y, y_test=y_test, d_obs=np.array(d_obs), xt_mean=xt_mean, xt_raw=xt_raw)
    out.update({"p_" + k: v for k, v in params.items()})
    for which in ("missing", "latent"):
        tf.reset_default_graph()
        tf.VARIABLE_OVERRIDES = [params[k] for k in names]
        np.random.seed(seed)
        model = bayesian_gp_lvm(y_train=y, num_latent_dims=q, num_inducing_points=m)
        vs = tf.get_collection("variables")
        assert len(vs) == len(names) and all(tuple(v.shape) == params[k].shape for v, k in zip(vs, names))
        if which == "missing":
            grads = torch.autograd.grad(model.objective, vs)
            out.update(objective=_np(model.objective), ard_weights=_np(model.ard_weights), noise_precision=_np(model.noise_precision))
            out.update({"g_" + k: _np(g) for k, g in zip(names, grads)})
        tf.VARIABLE_OVERRIDES = [None] * len(names) + [xt_mean, xt_raw]
        if which == "missing":
            lb, xm, xc, pm, pc = model.predict_missing_data(y_test=y_test[:, :d_obs])
            out.update(missing_predicted_mean=_np(pm), missing_predicted_covar=_np(pc))
        else:
            lb, xm, xc, tll = model.predict_new_latent_variables(y_test=y_test)
            out.update(latent_test_log_likelihood=_np(tll))
        tf.VARIABLE_OVERRIDES = None
        g = torch.autograd.grad(lb, tf.get_collection("variables")[-2:])
        out.update({which + "_lower_bound": _np(lb), which + "_g_xt_mean": _np(g[0]), which + "_g_xt_raw": _np(g[1])})
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote %-14s objective = %.15g" % (name, float(out["objective"])))


def main():
    tf = ref_env.activate()
    os.makedirs(OUT, exist_ok=True)
    kernel_fixture(tf, "kernel_b1", batch=1)
    kernel_fixture(tf, "kernel_b7", batch=7)
    dp_fixture(tf, "dp_n10_t20", d=10, t=20, mask_size=1, seed=1, alpha_prior=(1.05, 0.93))
    dp_fixture(tf, "dp_n12_t5_mask3", d=12, t=5, mask_size=3, seed=2, alpha_prior=(1.0, 1.0))
    # (N, D, Q, M, T): reference unit-test shapes (dpgplvm_unitttests.py:34-38, :143-147, :252-256) and
    # small versions of the BASELINE configs; both formulations at UNEQUAL atoms.
    shapes = {
        "unit": (50, 10, 3, 25, 8, 1, False), "t1": (50, 5, 3, 25, 1, 1, False), "d2t1": (5, 3, 1, 3, 1, 1, False),
        "c1": (100, 10, 2, 25, 8, 1, True), "q10": (60, 22, 10, 20, 5, 1, False),
        "mask3": (30, 12, 3, 10, 4, 3, False), "c3s": (90, 12, 10, 30, 6, 3, True),
    }
    for i, (key, (n, d, q, m, t, mask, zfx)) in enumerate(shapes.items()):
        for mode in ("t", "d"):
            model_fixture(tf, "%s_%s" % (mode, key), mode, n, d, q, m, t, mask_size=mask, seed=10 + i,
                          z_from_x=zfx, alpha_prior=(1.0, 1.0) if key != "unit" else (1.1, 0.9))
    # the reference's own initialisation (PCA, permutation subset, equal atoms): T-mode == D-mode there
    # (dpgplvm_unitttests.py:547-548).  scipy >= 1.12 starts ARPACK from an unseeded vector, so two PCA
    # calls differ at the 1e-8 level; the D-mode model is therefore evaluated AT the T-mode model's
    # initial variables rather than re-initialised.
    a, p_init = model_fixture(tf, "t_init", "t", 60, 12, 10, 25, 6, seed=3, override=False)
    b, _ = model_fixture(tf, "d_init", "d", 60, 12, 10, 25, 6, seed=3, override=True, params=p_init)
    assert abs(a - b) < 1e-11 * abs(a), (a, b)
    prediction_fixture(tf, "pred_d_small", n=40, d=8, q=3, m=12, t=4, n_test=7, d_obs=5, seed=21)
    bgplvm_fixture(tf, "bgplvm_q4", n=60, d=9, q=4, m=15, n_test=6, d_obs=5, seed=31)
    prediction_fixture(tf, "pred_d_q10", n=70, d=14, q=10, m=20, t=5, n_test=11, d_obs=9, seed=22)


if __name__ == "__main__":
    sys.exit(main())
