"""DP-GP-LVM (reference src/models/dp_gp_lvm.py): `dp_gp_lvm` ("D-mode", :22-510, hyper-parameters mixed
by phi, one kernel per output dimension) and `dp_gp_lvm_t` ("T-mode", :513-1029, one kernel per DP atom,
bound terms mixed by phi).  Same factory signatures, defaults, assertions and accessors.

What changes: the TensorFlow-1 graph (psi statistics materialised as [B,N,M,M,Q] / [D,N,N] tensors,
tf.cholesky chain, tf.gradients) is replaced by the streamed sufficient-statistics form evaluated by the
CUDA kernels behind include/dpgp.h:

    stats_fwd (local rows)  ->  [all-reduce]  ->  bound (M x M chain fwd+bwd)  ->  stats_bwd  ->  [all-reduce]

wrapped in one torch.autograd.Function so that `model.objective` is a differentiable torch scalar: callers
drive it with torch.optim.Adam exactly as the reference's scripts drive tf.train.AdamOptimizer
(test/synthetic_data_hard_test.py:143-155).  The O(D T) pieces (softplus / softmax, DP objective,
log-normal hyper-prior, phi-mixing in D-mode) are torch ops on the same device.

Data parallelism (not in the reference): each rank passes ITS rows of y_train and a process group; the
variational parameters of q(X) stay sharded, everything else is replicated, and the two all-reduces above
are the only communication.
"""
import numpy as np
import torch

from .. import engine as _engine
from ..distributions.log_normal import log_pdf as log_normal_log_pdf
from ..kernels.interfaces.kernel import KernelHyperparameters
from ..kernels.rbf_kernel import k_ard_rbf
from ..utils.constants import (DP_DEFAULT_ALPHA_PRIOR_PARAMS, DP_DEFAULT_TRUNCATION_LEVEL, GP_INIT_ALPHA, GP_INIT_BETA,
                               GP_INIT_GAMMA, GP_LVM_DEFAULT_LATENT_DIMENSIONS, GP_LVM_DEFAULT_NUM_INDUCING_POINTS)
from ..utils.expressions import principal_component_analysis as pca
from ..utils.types import TORCH_DTYPE, create_positive_variable
from .dirichlet_process import dirichlet_process
from .interfaces.trainable import Trainable

# Test hook: tests/test_distributed_cpu.py injects an oracle-backed engine here to exercise the N-sharding /
# all-reduce logic of this module on CPU (gloo).  The product never sets it: the default is the CUDA engine,
# which raises without a GPU.
ENGINE_FACTORY = None

PARAM_ORDER = ("x_mean", "x_var_raw", "x_u", "phi_logits", "gamma1_raw", "gamma2_raw", "w1_raw", "w2_raw",
               "gamma_atoms_raw", "alpha_atoms_raw", "beta_atoms_raw")


class _BoundFunction(torch.autograd.Function):
    """gp = f_hat - KL(q(X)||p(X)) and its gradient w.r.t. (x_mean, s, x_u, gamma [B,Q], alpha [B], beta [B], phi)."""

    @staticmethod
    def forward(ctx, eng, y, n_total, group, x_mean, s, x_u, gamma, alpha, beta, phi, psi2_hook=None, want_grad=True):
        tensors = [t.detach().contiguous() for t in (x_mean, s, x_u, gamma, alpha, beta)]
        mu_, s_, z_, g_, a_, b_ = tensors
        phi_ = None if phi is None else phi.detach().contiguous()
        # want_grad = torch.is_grad_enabled() at the call site (grad mode is always off inside Function.forward): a value-only
        # read under torch.no_grad() skips the statistics backward, which is two thirds of an evaluation
        need_grad = want_grad and any(t is not None and t.requires_grad for t in (x_mean, s, x_u, gamma, alpha, beta, phi))
        stats = eng.stats_fwd(mu_, s_, y, z_, g_, a_)
        if group is not None:
            torch.distributed.all_reduce(stats, group=group)
        gp, dstats, dz, dgamma, dalpha, dbeta, dphi = eng.bound(n_total, stats, z_, g_, a_, b_, phi_)
        if psi2_hook is not None:
            # extra term that is a function of Psi2 only (prediction.py: the reference's broadcast quirk): value and
            # cotangent are injected between the M x M chain and the statistics backward
            extra, dpsi2 = psi2_hook(eng.split_stats(stats)[0])
            gp = gp + extra.reshape(1)
            dstats[:dpsi2.numel()] += dpsi2.reshape(-1)
        if need_grad:
            dmu, ds, dz_s, dg_s, da_s = eng.stats_bwd(mu_, s_, y, z_, g_, a_, dstats)
            small = torch.cat([dz_s.reshape(-1), dg_s.reshape(-1), da_s.reshape(-1)])
            if group is not None:
                torch.distributed.all_reduce(small, group=group)
            nz, ng = dz.numel(), dgamma.numel()
            dz = dz + small[:nz].view_as(dz)
            dgamma = dgamma + small[nz:nz + ng].view_as(dgamma)
            dalpha = dalpha + small[nz + ng:].view_as(dalpha)
            ctx.save_for_backward(dmu, ds, dz, dgamma, dalpha, dbeta, dphi if dphi is not None else torch.empty(0, device=dz.device))
            ctx.has_phi = dphi is not None
            ctx.alpha_shape = alpha.shape
            ctx.beta_shape = beta.shape
        ctx.stats = stats
        return gp.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        dmu, ds, dz, dgamma, dalpha, dbeta, dphi = ctx.saved_tensors
        g = grad_out
        return (None, None, None, None, g * dmu, g * ds, g * dz, g * dgamma, (g * dalpha).view(ctx.alpha_shape),
                (g * dbeta).view(ctx.beta_shape), (g * dphi) if ctx.has_phi else None, None, None)


class _ObjectiveFunction(torch.autograd.Function):
    """objective = dp.objective - (f_hat - KL) - hyper-prior (dp_gp_lvm.py:148-154 / :670-676) and its gradient w.r.t. the
    eleven raw variables.  Same hot path as _BoundFunction; the N-independent part (softplus / softmax, the DP objective,
    the hyper-prior, the D-mode mixtures and their chain rule) runs as the two fused kernels behind dpgp_small_fwd /
    dpgp_small_bwd instead of ~300 torch ops per evaluation -- at the reference's own problem sizes those launches were
    most of a training iteration."""

    @staticmethod
    def forward(ctx, eng, y, n_total, group, meta, x_mean, x_var_raw, x_u, logits, g1, g2, w1, w2, ga, aa, ba):
        trunc, mask, prior, mode, want_grad = meta
        det = lambda t: t.detach().contiguous()
        raw = {"logits": det(logits), "gamma1_raw": det(g1) if trunc > 1 else None, "gamma2_raw": det(g2) if trunc > 1 else None,
               "w1_raw": det(w1), "w2_raw": det(w2), "gamma_atoms_raw": det(ga), "alpha_atoms_raw": det(aa), "beta_atoms_raw": det(ba)}
        mu_, z_, xr_ = det(x_mean), det(x_u), det(x_var_raw)
        s_ = torch.nn.functional.softplus(xr_, beta=1.0, threshold=1.0e9)
        phi, gam, alp, bet, scal = eng.small_fwd(raw, trunc, mask, prior)
        stats = eng.stats_fwd(mu_, s_, y, z_, gam, alp)
        if group is not None:
            torch.distributed.all_reduce(stats, group=group)
        gp, dstats, dz, dgamma, dalpha, dbeta, dphi = eng.bound(n_total, stats, z_, gam, alp, bet, phi if mode == "t" else None)
        obj = (scal[0] - scal[1]) - gp[0]
        leaves = (x_mean, x_var_raw, x_u, logits, g1, g2, w1, w2, ga, aa, ba)
        if want_grad and any(t.requires_grad for t in leaves):
            dmu, ds, dz_s, dg_s, da_s = eng.stats_bwd(mu_, s_, y, z_, gam, alp, dstats)
            small = torch.cat([dz_s.reshape(-1), dg_s.reshape(-1), da_s.reshape(-1)])
            if group is not None:
                torch.distributed.all_reduce(small, group=group)
            nz, ng = dz.numel(), dgamma.numel()
            small[:nz] += dz.reshape(-1); small[nz:nz + ng] += dgamma.reshape(-1); small[nz + ng:] += dalpha.reshape(-1)
            dz_t = small[:nz].view_as(dz)
            # raw-variable gradients of the small variables, packed: logits | g1 | g2 | w1 | w2 | gamma atoms | alpha | beta
            sizes = [logits.numel(), g1.numel(), g2.numel(), 1, 1, ga.numel(), aa.numel(), ba.numel()]
            packed = torch.empty(sum(sizes), dtype=torch.float64, device=dz.device)
            parts = torch.split(packed, sizes)
            names = ("dlogits", "dgamma1_raw", "dgamma2_raw", "dw1_raw", "dw2_raw", "dgamma_atoms_raw", "dalpha_atoms_raw", "dbeta_atoms_raw")
            out = {k: (v if v.numel() else None) for k, v in zip(names, parts)}
            eng.small_bwd(raw, trunc, mask, prior, phi, dphi, small[nz:nz + ng].view_as(dgamma), small[nz + ng:], dbeta, out)
            ds.mul_(torch.sigmoid(xr_))                         # chain of s = softplus(raw)
            ctx.save_for_backward(dmu, ds, dz_t, packed)
            ctx.sizes = sizes
            ctx.shapes = [t.shape for t in (logits, g1, g2, w1, w2, ga, aa, ba)]
        ctx.stats = stats
        return obj.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        dmu, ds, dz, packed = ctx.saved_tensors
        ng = -grad_out                                          # the bound enters the objective with a minus sign
        parts = torch.split(packed * grad_out, ctx.sizes)
        small = tuple(p.view(shp) for p, shp in zip(parts, ctx.shapes))
        return (None, None, None, None, None, ng * dmu, ng * ds, ng * dz) + small


def _build(y_train, num_latent_dims, num_inducing_points, truncation_level, alpha_prior_params, mask_size, mode,
           device, process_group, exp_variant, bwd_variant=0):
    num_samples, num_dimensions = np.shape(y_train)
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("dp_gp_lvm_b200 needs a CUDA device (no CPU fallback)")
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    dist_on = process_group is not None
    world = torch.distributed.get_world_size(process_group) if dist_on else 1
    rank = torch.distributed.get_rank(process_group) if dist_on else 0

    y_dev = torch.as_tensor(np.ascontiguousarray(y_train, dtype=np.float64), device=device)
    n_total = num_samples
    if dist_on:
        cnt = torch.tensor([num_samples], dtype=torch.int64, device=device)
        torch.distributed.all_reduce(cnt, group=process_group)
        n_total = int(cnt.item())

    # Fit latent means using PCA and the inducing inputs as a subset of those with a little noise.
    if dist_on:
        x_init = _distributed_pca(y_dev, num_latent_dims, n_total, process_group)
    else:
        x_init = pca(np.asarray(y_train, dtype=np.float64), num_latent_dimensions=num_latent_dims)
    x_mean = torch.tensor(x_init, dtype=TORCH_DTYPE, device=device, requires_grad=True)          # [N x Q]
    # q(X) variances: positive variable initialised to 1.0 (dp_gp_lvm.py:67-69 / :568-570)
    x_var = create_positive_variable(initial_value=1.0, shape=(num_samples, num_latent_dims), device=device)
    x_u_init = np.random.permutation(x_init)[:num_inducing_points] + \
        np.random.normal(loc=0.0, scale=0.01, size=(num_inducing_points, num_latent_dims))
    x_u = torch.tensor(x_u_init, dtype=TORCH_DTYPE, device=device, requires_grad=True)           # [M x Q]
    if dist_on:
        with torch.no_grad():
            torch.distributed.broadcast(x_u, src=torch.distributed.get_global_rank(process_group, 0), group=process_group)

    dp_model = dirichlet_process(num_samples=num_dimensions, alpha_prior_params=alpha_prior_params,
                                 truncation_level=truncation_level, mask_size=mask_size, device=device)
    gamma_atoms = create_positive_variable(initial_value=GP_INIT_GAMMA, shape=(truncation_level, num_latent_dims), device=device)
    sig_var_atoms = create_positive_variable(initial_value=GP_INIT_ALPHA, shape=(truncation_level, 1), device=device)
    beta_atoms = create_positive_variable(initial_value=GP_INIT_BETA, shape=(truncation_level, 1), device=device)
    if dist_on:
        with torch.no_grad():
            src = torch.distributed.get_global_rank(process_group, 0)
            for _, v in dp_model.variables:
                torch.distributed.broadcast(v, src=src, group=process_group)

    batch = truncation_level if mode == "t" else num_dimensions
    make_engine = ENGINE_FACTORY if ENGINE_FACTORY is not None else _engine.BoundEngine
    eng = make_engine(num_samples, num_dimensions, num_latent_dims, num_inducing_points, batch,
                      _engine.MODE_T if mode == "t" else _engine.MODE_D, device=device, exp_variant=exp_variant,
                      bwd_variant=bwd_variant)

    def hyper():
        """Kernel-batch hyper-parameters: the atoms (T-mode, :608) or their phi-mixtures (D-mode, :100-102)."""
        if mode == "t":
            return gamma_atoms.value, sig_var_atoms.value, beta_atoms.value
        phi = dp_model.assignments
        return phi @ gamma_atoms.value, phi @ sig_var_atoms.value, phi @ beta_atoms.value

    kernel = k_ard_rbf(gamma=lambda: hyper()[0], alpha=lambda: hyper()[1], beta=lambda: hyper()[2], device=device)

    def hyperprior():
        return torch.sum(log_normal_log_pdf(gamma_atoms.value)) + torch.sum(log_normal_log_pdf(sig_var_atoms.value)) + \
            torch.sum(log_normal_log_pdf(beta_atoms.value))

    state = {"last_stats": None, "fused_small": hasattr(eng, "small_fwd")}
    meta = (truncation_level, mask_size, (float(alpha_prior_params[0]), float(alpha_prior_params[1])), mode)

    def objective_value():
        if state["fused_small"]:
            dpv = dict(dp_model.variables)
            return _ObjectiveFunction.apply(eng, y_dev, n_total, process_group, meta + (torch.is_grad_enabled(),), x_mean, x_var.raw, x_u, dpv["phi_logits"],
                                            dpv["gamma1_raw"], dpv["gamma2_raw"], dpv["w1_raw"], dpv["w2_raw"],
                                            gamma_atoms.raw, sig_var_atoms.raw, beta_atoms.raw)
        phi = dp_model.assignments
        if mode == "t":
            gam, alp, bet = gamma_atoms.value, sig_var_atoms.value, beta_atoms.value
            gp_elbo = _BoundFunction.apply(eng, y_dev, n_total, process_group, x_mean, x_var.value, x_u, gam,
                                           alp.reshape(-1), bet.reshape(-1), phi, None, torch.is_grad_enabled())
        else:
            gam, alp, bet = phi @ gamma_atoms.value, phi @ sig_var_atoms.value, phi @ beta_atoms.value
            gp_elbo = _BoundFunction.apply(eng, y_dev, n_total, process_group, x_mean, x_var.value, x_u, gam,
                                           alp.reshape(-1), bet.reshape(-1), None, None, torch.is_grad_enabled())
        # objective = dp.objective - (f_hat - KL) - hyper-prior   (dp_gp_lvm.py:148-154 / :670-676)
        return dp_model.objective_at(phi) - gp_elbo - hyperprior()

    def fused_adam_iteration(params, ms, vs, step, objective_out, lr, beta1, beta2, eps):
        """One iteration of `AdamOptimizer(...).minimize(model)` (objective, all gradients, TF-1 Adam update of every variable)
        straight through the C ABI: no autograd graph and none of the ~14 elementwise torch launches that glue the pieces
        together in `_ObjectiveFunction` (sign flips, the softplus chain, cat / add of the small gradients, the step counter) --
        at the reference's own problem sizes those were 14 of 42 launches of an iteration.  Same arithmetic, same order.
        Returns False (nothing done) when the fused small-variable kernels are off or `params` is not the model's own list."""
        if not state["fused_small"] or not hasattr(eng, "train_tail"):
            return False
        dpv = dict(dp_model.variables)
        leaves = (x_mean, x_var.raw, x_u, dpv["phi_logits"], dpv["gamma1_raw"], dpv["gamma2_raw"], dpv["w1_raw"], dpv["w2_raw"],
                  gamma_atoms.raw, sig_var_atoms.raw, beta_atoms.raw)
        if len(params) != len(leaves) or any(a is not b for a, b in zip(params, leaves)):
            return False
        trunc, mask, prior, md = meta
        with torch.no_grad():
            det = lambda t: t.detach().contiguous()
            logits, g1, g2, w1, w2, ga, aa, ba = [det(t) for t in leaves[3:]]
            raw = {"logits": logits, "gamma1_raw": g1 if trunc > 1 else None, "gamma2_raw": g2 if trunc > 1 else None, "w1_raw": w1,
                   "w2_raw": w2, "gamma_atoms_raw": ga, "alpha_atoms_raw": aa, "beta_atoms_raw": ba}
            mu_, xr_, z_ = det(x_mean), det(x_var.raw), det(x_u)
            s_ = torch.nn.functional.softplus(xr_, beta=1.0, threshold=1.0e9)
            phi, gam, alp, bet, scal = eng.small_fwd(raw, trunc, mask, prior)
            stats = eng.stats_fwd(mu_, s_, y_dev, z_, gam, alp)
            if process_group is not None:
                torch.distributed.all_reduce(stats, group=process_group)
            nsmall = z_.numel() + gam.numel() + alp.numel()
            bsmall = torch.empty(nsmall, dtype=TORCH_DTYPE, device=z_.device)
            small = torch.empty(nsmall, dtype=TORCH_DTYPE, device=z_.device)
            gp, dstats, _, _, _, dbeta, dphi = eng.bound(n_total, stats, z_, gam, alp, bet, phi if md == "t" else None, small_out=bsmall)
            dmu, ds, dz_t, dg_t, da_t = eng.stats_bwd(mu_, s_, y_dev, z_, gam, alp, dstats, small_out=small)
            if process_group is not None:
                torch.distributed.all_reduce(small, group=process_group)
            small += bsmall                                     # [dz | dgamma | dalpha]: statistics part + M x M chain part
            sizes = [logits.numel(), g1.numel(), g2.numel(), 1, 1, ga.numel(), aa.numel(), ba.numel()]
            packed = torch.empty(sum(sizes), dtype=TORCH_DTYPE, device=z_.device)
            parts = torch.split(packed, sizes)
            names = ("dlogits", "dgamma1_raw", "dgamma2_raw", "dw1_raw", "dw2_raw", "dgamma_atoms_raw", "dalpha_atoms_raw", "dbeta_atoms_raw")
            out = {k: (v if v.numel() else None) for k, v in zip(names, parts)}
            eng.small_bwd(raw, trunc, mask, prior, phi, dphi, dg_t, da_t, dbeta, out)
            # gradients of the objective: -(d gp) for q(X) and the inducing inputs (x_var through the softplus chain), the packed raw
            # gradients of dpgp_small_bwd as they are
            grads = [dmu, ds, dz_t] + list(parts)
            scales = [-1.0, -1.0, -1.0] + [1.0] * len(parts)
            raws = [None, xr_, None] + [None] * len(parts)
            keep = [i for i, g in enumerate(grads) if g.numel() > 0]
            eng.train_tail(scal, gp, objective_out, [params[i].data for i in keep], [grads[i] for i in keep], [ms[i] for i in keep],
                           [vs[i] for i in keep], [scales[i] for i in keep], [raws[i] for i in keep], step, lr, beta1, beta2, eps)
            state["last_stats"] = stats
        return True

    def prediction_context():
        return {"device": device, "mode": mode, "engine": eng, "y_dev": y_dev, "x_mean": x_mean, "x_var": x_var, "x_u": x_u,
                "hyper": lambda: (hyper()[0], hyper()[1], hyper()[2]), "dp": dp_model, "n_total": n_total,
                "process_group": process_group, "num_latent_dims": num_latent_dims, "num_inducing_points": num_inducing_points,
                "num_dimensions": num_dimensions, "truncation_level": truncation_level}

    class DP_GP_LVM(Trainable):
        @property
        def objective(self):
            return objective_value()

        @staticmethod
        def predict_new_latent_variables(y_test, use_pca=False, reference_broadcast=True):
            """Reference dp_gp_lvm.py:234-311 (and :755-832): returns an object that unpacks to
            (prediction_lower_bound, x_test_mean, x_test_covar, test_log_likelihood) and is itself a Trainable whose
            variables are q(X*) (models/prediction.py)."""
            from .prediction import LatentPrediction
            num_test_points, test_dims = np.shape(y_test)
            assert test_dims == num_dimensions, \
                'Observed dimensionality for prediction must be equal to the dimensionality of the training data.'
            return LatentPrediction(prediction_context(), np.asarray(y_test), test_dims, use_pca, reference_broadcast)

        @staticmethod
        def predict_missing_data(y_test, use_pca=False, reference_broadcast=True):
            """Reference dp_gp_lvm.py:313-500 (and :834-1018): y_test holds the first Do < D dimensions; unpacks to
            (missing_data_lower_bound, x_test_mean, x_test_covar, predicted_mean [N* x Du], predicted_covar [Du x N* x N*])."""
            from .prediction import MissingDataPrediction
            num_test_points, num_observed_dims = np.shape(y_test)
            assert num_observed_dims < num_dimensions, \
                'Observed dimensionality for missing data scenario must be less than total ' \
                'dimensionality of training data.'
            return MissingDataPrediction(prediction_context(), np.asarray(y_test), use_pca, reference_broadcast)

        @property
        def assignments(self):
            return dp_model.assignments

        @property
        def dp(self):
            return dp_model

        @property
        def dp_atoms(self):
            return gamma_atoms.value, sig_var_atoms.value, beta_atoms.value

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
            """(mean [N x Q], covariance [N x Q x Q] diagonal-embedded), as the reference."""
            return x_mean, torch.diag_embed(x_var.value)

        # ------------------------------------------------------------------ additions over the reference
        @property
        def variables(self):
            """Trainable leaves by name, in the reference's tf.Variable creation order."""
            dpv = dict(dp_model.variables)
            return {"x_mean": x_mean, "x_var_raw": x_var.raw, "x_u": x_u, "phi_logits": dpv["phi_logits"],
                    "gamma1_raw": dpv["gamma1_raw"], "gamma2_raw": dpv["gamma2_raw"], "w1_raw": dpv["w1_raw"],
                    "w2_raw": dpv["w2_raw"], "gamma_atoms_raw": gamma_atoms.raw, "alpha_atoms_raw": sig_var_atoms.raw,
                    "beta_atoms_raw": beta_atoms.raw}

        def parameters(self):
            return [self.variables[k] for k in PARAM_ORDER]

        def load_variables(self, values):
            """Overwrite leaves in place from a {name: array} dict (tests, checkpoints)."""
            with torch.no_grad():
                for k, v in values.items():
                    t = self.variables[k]
                    t.copy_(torch.as_tensor(np.asarray(v, dtype=np.float64), device=t.device).reshape(t.shape))

        def value_and_grad(self):
            """(objective as float, {name: gradient numpy array}); checks the device-side Cholesky flags."""
            params = self.parameters()
            for p in params:
                p.grad = None
            obj = objective_value()
            grads = torch.autograd.grad(obj, params, allow_unused=True)
            eng.check()
            return float(obj.item()), {k: (torch.zeros_like(p) if g is None else g).detach().cpu().numpy()
                                        for k, p, g in zip(PARAM_ORDER, params, grads)}

        @property
        def fused_small(self):
            """True: the N-independent part of the objective runs as two fused kernels (dpgp_small_fwd / _bwd); False: as
            torch ops (the first implementation, kept as a cross-check)."""
            return state["fused_small"]

        @fused_small.setter
        def fused_small(self, value):
            state["fused_small"] = bool(value) and hasattr(eng, "small_fwd")

        @property
        def engine(self):
            return eng

        @staticmethod
        def fused_adam_iteration(params, ms, vs, step, objective_out, lr, beta1, beta2, eps):
            return fused_adam_iteration(params, ms, vs, step, objective_out, lr, beta1, beta2, eps)

        @property
        def y_train_device(self):
            """This rank's rows of y_train on the device (the reference bakes y_train into the graph as a constant)."""
            return y_dev

        @property
        def num_samples_total(self):
            return n_total

    return DP_GP_LVM()


def _distributed_pca(y_dev, q, n_total, group):
    """PCA initialisation for row-sharded Y: eigenvectors of the all-reduced D x D matrix Y^T Y."""
    yty = y_dev.T @ y_dev
    torch.distributed.all_reduce(yty, group=group)
    w, v = torch.linalg.eigh(yty)
    idx = torch.argsort(w, descending=True)[:q]
    x0 = (y_dev @ v[:, idx]) / torch.sqrt(torch.clamp(w[idx], min=1e-300))
    s1 = x0.sum(0); s2 = (x0 ** 2).sum(0)
    both = torch.stack([s1, s2]); torch.distributed.all_reduce(both, group=group)
    var = (both[1] - both[0] ** 2 / n_total) / (n_total - 1)
    return (x0 / torch.sqrt(var).mean()).cpu().numpy()


def dp_gp_lvm(y_train,
              num_latent_dims=GP_LVM_DEFAULT_LATENT_DIMENSIONS,
              num_inducing_points=GP_LVM_DEFAULT_NUM_INDUCING_POINTS,
              truncation_level=DP_DEFAULT_TRUNCATION_LEVEL,
              alpha_prior_params=DP_DEFAULT_ALPHA_PRIOR_PARAMS,
              mask_size=1, device=None, process_group=None, exp_variant=0, bwd_variant=0):
    """D-mode DP-GP-LVM, reference src/models/dp_gp_lvm.py:22-154 (same arguments; `device`,
    `process_group`, `exp_variant`, `bwd_variant` are additions).  Relies on the caller to seed numpy, as the reference."""
    num_samples, num_dimensions = np.shape(y_train)
    assert 0 < num_latent_dims <= num_dimensions, \
        'Number of latent dimensions must be postive and less than the dimensionality of the observed data.'
    assert 0 < num_inducing_points <= num_samples, \
        'Number of inducing points must be positive and less than the number of observations in the observed data.'
    assert 0 < truncation_level <= min(num_samples, num_dimensions), \
        'The truncation level must be positive and less than the dimensionality of the observed data and ' \
        'less than the number of observations.'
    return _build(y_train, num_latent_dims, num_inducing_points, truncation_level, alpha_prior_params, mask_size, "d",
                  device, process_group, exp_variant, bwd_variant)


def dp_gp_lvm_t(y_train,
                num_latent_dims=GP_LVM_DEFAULT_LATENT_DIMENSIONS,
                num_inducing_points=GP_LVM_DEFAULT_NUM_INDUCING_POINTS,
                truncation_level=DP_DEFAULT_TRUNCATION_LEVEL,
                alpha_prior_params=DP_DEFAULT_ALPHA_PRIOR_PARAMS,
                mask_size=1,
                seed=0, device=None, process_group=None, exp_variant=0, bwd_variant=0):
    """T-mode DP-GP-LVM, reference src/models/dp_gp_lvm.py:513-676 (same arguments and assertions; seeds
    numpy inside the factory as the reference does, :559-560)."""
    assert isinstance(y_train, np.ndarray), 'Training data must be provided as a numpy array.'
    num_samples, num_dimensions = np.shape(y_train)
    assert isinstance(num_latent_dims, int), 'Number of latent dimensions must be an integer.'
    assert 0 < num_latent_dims < num_dimensions, \
        'Number of latent dimensions must be postive and less than the dimensionality of the observed data.'
    assert isinstance(num_inducing_points, int), 'Number of inducing points must be an integer.'
    assert 0 < num_inducing_points <= num_samples, \
        'Number of inducing points must be positive and less than or equal to the number of observations in the ' \
        'observed data.'
    assert isinstance(truncation_level, int), 'The truncation level must be an integer.'
    assert 0 < truncation_level <= min(num_samples, num_dimensions), \
        'The truncation level must be positive and less than or equal to the dimensionality of the observed data and ' \
        'less than or equal to the number of observations.'
    assert isinstance(seed, int) and seed >= 0, 'Seed must be a 32-bit unsigned integer, i.e., 0 <= seed <= 2^32 - 1.'
    np.random.seed(seed=seed)
    return _build(y_train, num_latent_dims, num_inducing_points, truncation_level, alpha_prior_params, mask_size, "t",
                  device, process_group, exp_variant, bwd_variant)
