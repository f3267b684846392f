"""Bayesian GP-LVM (reference src/models/gaussian_process.py:132-548): the B = 1 special case of the streamed bound.

The reference's `bayesian_gp_lvm` has one RBF-ARD kernel shared by all D output dimensions; its f_hat (:247-251)
    1/2 N D (log beta - log 2 pi) - D log|L_A| + 1/2 D beta (tr H - psi0) + 1/2 beta^2 tr(C^T C Y Y^T) - 1/2 beta tr(Y Y^T)
is the T-mode DP-GP-LVM bound with a single cluster and phi = 1, so it runs on the same CUDA kernels (BoundEngine in
T-mode with B = 1).  Same factory signature, assertions, initialisation (q(X) variances 0.5, :223) and accessors;
the objective is -(ELBO + kernel.prior_log_likelihood) (:260).  The Monte-Carlo branch (num_latent_samples > 0, which
needs tensorflow_probability sampling) is outside the hot path and raises NotImplementedError.  The other models of
that file (gp_regression, gp_lvm, manifold_relevance_determination) are out of scope."""
import numpy as np
import torch

from .. import engine as _engine
from ..kernels.interfaces.kernel import KernelHyperparameters
from ..kernels.rbf_kernel import k_ard_rbf
from ..utils.constants import (GP_INIT_ALPHA, GP_INIT_BETA, GP_INIT_GAMMA, GP_LVM_DEFAULT_LATENT_DIMENSIONS,
                               GP_LVM_DEFAULT_NUM_INDUCING_POINTS, MAX_MC_SAMPLES)
from ..utils.expressions import principal_component_analysis as pca
from ..utils.types import TORCH_DTYPE, create_positive_variable
from .interfaces.trainable import Trainable

BGPLVM_PARAM_ORDER = ("gamma_raw", "alpha_raw", "beta_raw", "x_mean", "x_u", "x_var_raw")     # tf.Variable creation order


class _Ones:
    """Stand-in for the DP of a one-cluster model: every dimension is assigned to the single kernel."""

    def __init__(self, d, device):
        self.assignments = torch.ones(d, 1, dtype=TORCH_DTYPE, device=device)


def bayesian_gp_lvm(y_train, kernel=None, num_latent_dims=GP_LVM_DEFAULT_LATENT_DIMENSIONS,
                    num_inducing_points=GP_LVM_DEFAULT_NUM_INDUCING_POINTS, num_latent_samples=0, device=None,
                    process_group=None):
    from .dp_gp_lvm import ENGINE_FACTORY, _BoundFunction, _distributed_pca
    num_samples, num_dimensions = np.shape(y_train)
    assert isinstance(num_latent_dims, int), 'Number of latent dimensions must be an integer.'
    assert 0 < num_latent_dims < num_dimensions, \
        'Number of latent dimensions must be postive and less than the dimensionality of the observed data.'
    assert isinstance(num_inducing_points, int), 'Number of inducing points must be an integer.'
    assert 0 < num_inducing_points < num_samples, \
        'Number of inducing points must be positive and less than the number of observations in the observed data.'
    assert isinstance(num_latent_samples, int), 'Number of latent space samples must be an integer.'
    assert 0 <= num_latent_samples < MAX_MC_SAMPLES, \
        'Number of latent space samples must be positive and less than {}.'.format(MAX_MC_SAMPLES)
    if num_latent_samples:
        raise NotImplementedError("the Monte-Carlo (SVI) psi statistics of gaussian_process.py:196-218 are not on the hot path")
    if kernel is not None:
        raise NotImplementedError("pass kernel=None: the model owns its RBF-ARD kernel variables (gaussian_process.py:176-186)")
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("dp_gp_lvm_b200 needs a CUDA device (no CPU fallback)")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    dist_on = process_group is not None
    y_dev = torch.as_tensor(np.ascontiguousarray(y_train, dtype=np.float64), device=device)
    n_total = num_samples
    if dist_on:
        cnt = torch.tensor([num_samples], dtype=torch.int64, device=device)
        torch.distributed.all_reduce(cnt, group=process_group)
        n_total = int(cnt.item())

    batch_size = 1
    gamma = create_positive_variable(initial_value=GP_INIT_GAMMA, shape=(batch_size, num_latent_dims), device=device)
    alpha = create_positive_variable(initial_value=GP_INIT_ALPHA, shape=(batch_size, 1), device=device)
    beta = create_positive_variable(initial_value=GP_INIT_BETA, shape=(batch_size, 1), device=device)
    kernel = k_ard_rbf(gamma=lambda: gamma.value, alpha=lambda: alpha.value, beta=lambda: beta.value, device=device)

    x_init = _distributed_pca(y_dev, num_latent_dims, n_total, process_group) if dist_on else \
        pca(np.asarray(y_train, dtype=np.float64), num_latent_dimensions=num_latent_dims)
    x_mean = torch.tensor(x_init, dtype=TORCH_DTYPE, device=device, requires_grad=True)
    x_u_init = np.random.permutation(x_init)[:num_inducing_points] + \
        np.random.normal(loc=0.0, scale=0.01, size=(num_inducing_points, num_latent_dims))
    x_u = torch.tensor(x_u_init, dtype=TORCH_DTYPE, device=device, requires_grad=True)
    if dist_on:
        with torch.no_grad():
            torch.distributed.broadcast(x_u, src=torch.distributed.get_global_rank(process_group, 0), group=process_group)
    x_var = create_positive_variable(initial_value=0.5, shape=(num_samples, num_latent_dims), device=device)

    make_engine = ENGINE_FACTORY if ENGINE_FACTORY is not None else _engine.BoundEngine
    eng = make_engine(num_samples, num_dimensions, num_latent_dims, num_inducing_points, batch_size, _engine.MODE_T, device=device)
    ones = _Ones(num_dimensions, device)

    def objective_value():
        elbo = _BoundFunction.apply(eng, y_dev, n_total, process_group, x_mean, x_var.value, x_u, gamma.value,
                                    alpha.value.reshape(-1), beta.value.reshape(-1), ones.assignments)
        return -(elbo + kernel.prior_log_likelihood)

    def prediction_context():
        return {"device": device, "mode": "t", "engine": eng, "y_dev": y_dev, "x_mean": x_mean, "x_var": x_var, "x_u": x_u,
                "hyper": lambda: (gamma.value, alpha.value, beta.value), "dp": ones, "n_total": n_total,
                "process_group": process_group, "num_latent_dims": num_latent_dims, "num_inducing_points": num_inducing_points,
                "num_dimensions": num_dimensions, "truncation_level": 1, "bgplvm": True}

    class BayesianGPLVM(Trainable):
        @property
        def kernel(self):
            return kernel

        @property
        def ard_weights(self):
            return kernel.hyperparameters[KernelHyperparameters.ARD_WEIGHTS]

        @property
        def signal_variance(self):
            return kernel.hyperparameters[KernelHyperparameters.SIGNAL_VARIANCE]

        @property
        def noise_precision(self):
            return kernel.noise_precision

        @property
        def inducing_input(self):
            return x_u

        @property
        def q_x(self):
            return x_mean, torch.diag_embed(x_var.value)

        @staticmethod
        def predict_new_latent_variables(y_test, use_pca=False):
            """gaussian_process.py:328-405; note the reference's `test_log_likelihood = f_hat_test - f_hat` (:403)."""
            from .prediction import LatentPrediction
            num_test_points, test_dims = np.shape(y_test)
            assert test_dims == num_dimensions, \
                'Observed dimensionality for prediction must be equal to the dimensionality of the training data.'
            return LatentPrediction(prediction_context(), np.asarray(y_test), test_dims, use_pca)

        @staticmethod
        def predict_missing_data(y_test, use_pca=False):
            """gaussian_process.py:407-537."""
            from .prediction import MissingDataPrediction
            num_test_points, num_observed_dims = np.shape(y_test)
            assert num_observed_dims < num_dimensions, \
                'Observed dimensionality for missing data scenario must be less than the total ' \
                'dimensionality of the training data.'
            return MissingDataPrediction(prediction_context(), np.asarray(y_test), use_pca)

        @property
        def objective(self):
            return objective_value()

        # ------------------------------------------------------------------ additions over the reference
        @property
        def variables(self):
            return {"gamma_raw": gamma.raw, "alpha_raw": alpha.raw, "beta_raw": beta.raw, "x_mean": x_mean, "x_u": x_u,
                    "x_var_raw": x_var.raw}

        def parameters(self):
            return [self.variables[k] for k in BGPLVM_PARAM_ORDER]

        def load_variables(self, values):
            with torch.no_grad():
                for k, v in values.items():
                    t = self.variables[k]
                    t.copy_(torch.as_tensor(np.asarray(v, dtype=np.float64), device=t.device).reshape(t.shape))

        def value_and_grad(self):
            params = self.parameters()
            obj = objective_value()
            grads = torch.autograd.grad(obj, params, allow_unused=True)
            eng.check()
            return float(obj.item()), {k: (torch.zeros_like(p) if g is None else g).detach().cpu().numpy()
                                        for k, p, g in zip(BGPLVM_PARAM_ORDER, params, grads)}

        @property
        def engine(self):
            return eng

        @property
        def y_train_device(self):
            return y_dev

    return BayesianGPLVM()
