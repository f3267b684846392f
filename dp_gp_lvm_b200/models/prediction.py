"""Prediction paths of the DP-GP-LVM (reference src/models/dp_gp_lvm.py:234-500 for the D-mode model,
:755-1018 for the T-mode copy of it, which raises NameError upstream at :797, :812, :849 and is FIXED here, not copied).

The reference builds, inside the training graph, a second bound over the test points whose only free variables are
q(X*) = N(x_test_mean, diag(x_test_covar)) -- created with trainable=False and optimised afterwards with a second
Adam over `get_prediction_variables()` (test/frey_faces_prediction.py:166-215).  Here the same object is a small
`Trainable` of its own:

    pred = model.predict_missing_data(y_test=y_test_observed)
    lower_bound, x_test_mean, x_test_covar, predicted_mean, predicted_covar = pred      # values now (5-tuple)
    op = AdamOptimizer(learning_rate).minimize(loss=pred)                               # loss = -lower bound
    for c in range(predict_iter): op.run()
    predicted_mean, predicted_covar = pred.predicted_mean, pred.predicted_covar          # values after optimisation

Streaming form.  f_hat_test (dp_gp_lvm.py:263-300, :371-410) is exactly the training bound (:113-145) evaluated on
(Y*, q(X*)) for the observed dimensions with the training K_uu, so it is one more BoundEngine over N* rows and its
gradients w.r.t. q(X*) come from the same CUDA kernels.  The predictive moments only need per-kernel M x M factors of
the TRAINING bound (dpgp_bound_factors): with  S = (K_uu + beta Psi2)^-1,  u_d = S Psi1^T y_d,
    predicted_mean[:, d]  = beta_d Psi1*_d u_d                                              (:412-424)
    predicted_covar[d]    = beta_d^2 u_d^T (Psi2*_d - Psi1*_d^T Psi1*_d) u_d  (every entry)
                            + (psi0*_d + 1/beta_d + tr((K_uu^-1 - S) Psi2*_d)) I            (:426-498)
T-mode: cluster t plays the role of the kernel, dimension d reads its column of P_t, and the per-cluster terms are
mixed with phi[d, t] -- at equal atoms this coincides with the D-mode formulas (tested).

Quirk preserved by default (`reference_broadcast=True`, D-mode only).  The reference's test bound subtracts
`psi_0_test` of shape [D x 1] from `tf.trace(...)` of shape [D] (dp_gp_lvm.py:294, :398-399; the training bound uses
keepdims=True, :140-142), so TensorFlow broadcasts to [D x D] and the term becomes
    1/2 sum_{i,j} beta_i (tr H*_j - psi0*_i)      instead of      1/2 sum_i beta_i (tr H*_i - psi0*_i).
The reference optimises q(X*) against that value, so a drop-in must reproduce it: the difference
    1/2 sum_j (sum_i beta_i - beta_j) tr(K_j^-1 Psi2*_j) - 1/2 (Do - 1) sum_i beta_i alpha_i N*
is added through `_BoundFunction`'s psi2 hook (value and cotangent).  `reference_broadcast=False` gives the bound of
the paper (and is what the fixed T-mode path always uses)."""
import numpy as np
import torch

from .. import engine as _engine
from ..utils.expressions import principal_component_analysis as pca
from ..utils.types import TORCH_DTYPE, create_positive_variable
from .interfaces.trainable import Trainable


def nearest_neighbour(x_train, x_test, chunk=65536):
    """Index of the nearest training row (L2) for every test row (reference src/utils/expressions.py:28-44); also
    returns the distances so that row-sharded callers can combine ranks."""
    best_d = torch.full((x_test.shape[0],), float("inf"), dtype=x_train.dtype, device=x_train.device)
    best_i = torch.zeros(x_test.shape[0], dtype=torch.int64, device=x_train.device)
    for lo in range(0, x_train.shape[0], chunk):
        d = torch.cdist(x_train[lo:lo + chunk], x_test)                  # [chunk x N*]
        dmin, imin = d.min(dim=0)
        upd = dmin < best_d
        best_d = torch.where(upd, dmin, best_d); best_i = torch.where(upd, imin + lo, best_i)
    return best_i, best_d


class _TestBound(Trainable):
    """Shared machinery: q(X*) variables, the observed-dimension test engine and the constant training terms."""

    def __init__(self, ctx, y_test, observed_dims, use_pca, reference_broadcast=True):
        self.ctx = ctx
        self.reference_broadcast = bool(reference_broadcast) and ctx["mode"] == "d"
        self.bgplvm = bool(ctx.get("bgplvm", False))
        dev = ctx["device"]
        self.mode = ctx["mode"]
        y_test = np.ascontiguousarray(y_test, dtype=np.float64)
        self.num_test, self.num_obs = y_test.shape
        self.obs = observed_dims
        self.y_test = torch.as_tensor(y_test, device=dev)
        q = ctx["num_latent_dims"]
        # ---- variables of q(X*), initialised as dp_gp_lvm.py:246-259 / :327-345
        if use_pca:
            init = torch.as_tensor(pca(y_test, num_latent_dimensions=q), dtype=TORCH_DTYPE, device=dev)
        else:
            idx, dist = nearest_neighbour(ctx["y_dev"][:, :self.num_obs].contiguous(), self.y_test)
            rows = ctx["x_mean"].detach()[idx]
            group = ctx["process_group"]
            if group is not None:                                          # rows are sharded: keep the globally nearest
                world = torch.distributed.get_world_size(group)
                ds = [torch.empty_like(dist) for _ in range(world)]; rs = [torch.empty_like(rows) for _ in range(world)]
                torch.distributed.all_gather(ds, dist, group=group); torch.distributed.all_gather(rs, rows, group=group)
                win = torch.stack(ds).argmin(dim=0)
                rows = torch.stack(rs)[win, torch.arange(rows.shape[0], device=dev)]
            noise = np.random.normal(scale=0.01, size=(self.num_test, q))
            init = rows + torch.as_tensor(noise, dtype=TORCH_DTYPE, device=dev)
        init = init.clone()
        if ctx["process_group"] is not None:
            # every rank optimises the SAME q(X*): the noise above comes from each rank's own numpy stream (which differs
            # as soon as the shards do), so rank 0's initial point is the one everybody starts from
            g = ctx["process_group"]
            torch.distributed.broadcast(init, src=torch.distributed.get_global_rank(g, 0), group=g)
        self.x_test_mean = init.requires_grad_(True)
        self.x_test_var = create_positive_variable(initial_value=1.0, shape=(self.num_test, q), is_trainable=False, device=dev)
        m = ctx["num_inducing_points"]
        batch_o = ctx["truncation_level"] if self.mode == "t" else self.num_obs
        self.eng_obs = _engine.BoundEngine(self.num_test, self.num_obs, q, m, batch_o,
                                           _engine.MODE_T if self.mode == "t" else _engine.MODE_D, device=dev)
        self.refresh()

    # ------------------------------------------------------------------------------------------------------
    def refresh(self):
        """(Re-)evaluates the terms that depend on the training variables only: f_hat - KL of the training set and the
        M x M factors of its bound.  Called at construction; call again if the model is trained further."""
        c = self.ctx
        with torch.no_grad():
            gam, alp, bet = [t.detach().contiguous() for t in c["hyper"]()]
            phi = c["dp"].assignments.detach().contiguous()
            eng = c["engine"]
            stats = eng.stats_fwd(c["x_mean"].detach().contiguous(), c["x_var"].value.detach().contiguous(), c["y_dev"],
                                  c["x_u"].detach().contiguous(), gam, alp.reshape(-1).contiguous())
            if c["process_group"] is not None:
                torch.distributed.all_reduce(stats, group=c["process_group"])
            out = eng.bound(c["n_total"], stats, c["x_u"].detach().contiguous(), gam, alp.reshape(-1).contiguous(),
                            bet.reshape(-1).contiguous(), phi if self.mode == "t" else None)
            self.gp_train = out[0].reshape(()).clone()
            self.kinv, self.sinv, self.u = eng.bound_factors()
            self.gamma, self.alpha, self.beta, self.phi = gam, alp.reshape(-1), bet.reshape(-1), phi
            self.x_u = c["x_u"].detach().contiguous()

    def _gp_test(self):
        """f_hat_test - KL(q(X*) || p(X*)) as a differentiable scalar (dp_gp_lvm.py:263-306 / :371-410)."""
        from .dp_gp_lvm import _BoundFunction
        if self.mode == "t":
            return _BoundFunction.apply(self.eng_obs, self.y_test, self.num_test, None, self.x_test_mean, self.x_test_var.value,
                                        self.x_u, self.gamma, self.alpha, self.beta, self.phi[:self.num_obs].contiguous())
        o = self.num_obs
        hook = None
        if self.reference_broadcast:
            beta, alpha, kinv = self.beta[:o], self.alpha[:o], self.kinv[:o]
            coef = 0.5 * (beta.sum() - beta)                                           # [Do]
            const = -0.5 * (o - 1) * (beta * alpha).sum() * float(self.num_test)

            def hook(psi2):
                tr = (kinv * psi2).sum(dim=(1, 2))                                     # tr(K_j^-1 Psi2*_j), K symmetric
                return (coef * tr).sum() + const, coef[:, None, None] * kinv
        return _BoundFunction.apply(self.eng_obs, self.y_test, self.num_test, None, self.x_test_mean, self.x_test_var.value,
                                    self.x_u, self.gamma[:o].contiguous(), self.alpha[:o].contiguous(),
                                    self.beta[:o].contiguous(), None, hook)

    @staticmethod
    def _kl(mean, var):
        return 0.5 * ((mean ** 2).sum() + (var - torch.log(var)).sum() - mean.numel())

    @property
    def test_log_likelihood(self):
        """DP-GP-LVM: f_hat_test - KL_test (dp_gp_lvm.py:309).  The reference's BGP-LVM instead returns
        f_hat_test - f_hat (gaussian_process.py:403); both are reproduced as they are."""
        if not self.bgplvm:
            return self._gp_test()
        c = self.ctx
        f_hat_test = self._gp_test() + self._kl(self.x_test_mean, self.x_test_var.value)
        f_hat = self.gp_train + self._kl(c["x_mean"].detach(), c["x_var"].value.detach()) if c["process_group"] is None \
            else self.gp_train + self._kl_train_distributed()
        return f_hat_test - f_hat

    def _kl_train_distributed(self):
        c = self.ctx
        m, v = c["x_mean"].detach(), c["x_var"].value.detach()
        parts = torch.stack([(m ** 2).sum() + (v - torch.log(v)).sum()])
        torch.distributed.all_reduce(parts, group=c["process_group"])
        return 0.5 * (parts[0] - c["n_total"] * m.shape[1])

    @property
    def lower_bound(self):
        """f_hat + f_hat_test - KL - KL_test (dp_gp_lvm.py:306 / :408)."""
        return self.gp_train + self._gp_test()

    @property
    def objective(self):
        return -self.lower_bound

    @property
    def x_test_covar(self):
        return torch.diag_embed(self.x_test_var.value)

    @property
    def variables(self):
        return {"x_test_mean": self.x_test_mean, "x_test_var_raw": self.x_test_var.raw}

    def parameters(self):
        """The reference's get_prediction_variables(): the non-trainable variables optimised at test time."""
        return [self.x_test_mean, self.x_test_var.raw]

    def load_variables(self, values):
        with torch.no_grad():
            for k, v in values.items():
                t = self.variables[k]
                t.copy_(torch.as_tensor(np.asarray(v, dtype=np.float64), device=t.device).reshape(t.shape))

    @property
    def engine(self):
        return self.eng_obs


class LatentPrediction(_TestBound):
    """`predict_new_latent_variables(y_test)`: all D dimensions observed (dp_gp_lvm.py:234-311)."""

    def __iter__(self):
        return iter((self.lower_bound, self.x_test_mean, self.x_test_covar, self.test_log_likelihood))


class MissingDataPrediction(_TestBound):
    """`predict_missing_data(y_test)`: the first Do dimensions observed, the remaining Du predicted (:313-500)."""

    def __init__(self, ctx, y_test, use_pca, reference_broadcast=True):
        super().__init__(ctx, y_test, y_test.shape[1], use_pca, reference_broadcast)
        d = ctx["num_dimensions"]
        self.num_unobs = d - self.num_obs
        batch_u = ctx["truncation_level"] if self.mode == "t" else self.num_unobs
        self.eng_unobs = _engine.BoundEngine(self.num_test, self.num_unobs, ctx["num_latent_dims"], ctx["num_inducing_points"],
                                             batch_u, _engine.MODE_T if self.mode == "t" else _engine.MODE_D, device=ctx["device"])
        self._y_dummy = torch.zeros(self.num_test, self.num_unobs, dtype=TORCH_DTYPE, device=ctx["device"])

    def _moments(self):
        with torch.no_grad():
            o = self.num_obs
            mu, s = self.x_test_mean.detach().contiguous(), self.x_test_var.value.detach().contiguous()
            if self.mode == "t":
                g, a, b = self.gamma, self.alpha, self.beta                         # per cluster
            else:
                g, a, b = self.gamma[o:].contiguous(), self.alpha[o:].contiguous(), self.beta[o:].contiguous()
            psi1 = self.eng_unobs.psi1(mu, s, self.x_u, g, a)                        # [Bu x N* x M]
            stats = self.eng_unobs.stats_fwd(mu, s, self._y_dummy, self.x_u, g, a)
            psi2 = self.eng_unobs.split_stats(stats)[0]                              # [Bu x M x M]
            nstar = float(self.num_test)
            if self.mode == "t":
                u = self.u[:, :, o:]                                                 # [T x M x Du]
                phi = self.phi[o:]                                                   # [Du x T]
                f = torch.einsum("tnm,tmd->tnd", psi1, u)                            # Psi1*_t u_td   [T x N* x Du]
                mean = torch.einsum("dt,t,tnd->nd", phi, b, f)
                quad = torch.einsum("tmd,tmk,tkd->td", u, psi2, u) - (f ** 2).sum(dim=1)          # [T x Du]
                yu_var = torch.einsum("dt,t,td->d", phi, b * b, quad)
                trace = ((self.kinv - self.sinv) * psi2).sum(dim=(1, 2))                           # [T]
                # DP-GP-LVM adds the trace term (dp_gp_lvm.py:478-498), BGP-LVM subtracts it (gaussian_process.py:516-534)
                diag = phi @ (a * nstar + 1.0 / b + (-trace if self.bgplvm else trace))            # [Du]
            else:
                u = self.u[o:, :, 0]                                                 # [Du x M]
                f = torch.einsum("dnm,dm->dn", psi1, u)                              # [Du x N*]
                mean = (b[:, None] * f).transpose(0, 1).contiguous()
                quad = torch.einsum("dm,dmk,dk->d", u, psi2, u) - (f ** 2).sum(dim=1)
                yu_var = b * b * quad
                trace = ((self.kinv[o:] - self.sinv[o:]) * psi2).sum(dim=(1, 2))
                diag = a * nstar + 1.0 / b + trace
            eye = torch.eye(self.num_test, dtype=TORCH_DTYPE, device=mean.device)
            covar = yu_var[:, None, None] + diag[:, None, None] * eye
            return mean, covar

    @property
    def predicted_mean(self):
        """[N* x Du]"""
        return self._moments()[0]

    @property
    def predicted_covar(self):
        """[Du x N* x N*]: one covariance per unobserved dimension, as the reference."""
        return self._moments()[1]

    def __iter__(self):
        mean, covar = self._moments()
        return iter((self.lower_bound, self.x_test_mean, self.x_test_covar, mean, covar))
