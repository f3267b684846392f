"""Truncated stick-breaking Dirichlet-process variational objective (reference
src/models/dirichlet_process.py:17-136): phi = softmax(logits) [num_samples x T] with optional
`mask_size` tying (:39-51), q(V) Beta parameters (:54-55), q(alpha) Gamma parameters (:58-59), the six
ELBO terms (:64-77) and objective = -ELBO (:80-88).  O(D T) work: evaluated with torch ops on the model's
device every time `.objective` is read (a TF-1 graph re-evaluates on every session.run).  The DP-GP-LVM models' own
`.objective` does not come through here: it evaluates the same terms and their closed-form gradients in the fused
kernels behind dpgp_small_fwd / dpgp_small_bwd (csrc/small.cuh); this module serves the stand-alone `dirichlet_process`
API, the accessors and the cross-check (`model.fused_small = False`).  Quirk kept for
parity: the q(Z) entropy counts all `num_samples` rows even when mask_size > 1 (dp_gp_lvm.py:584-588)."""
import math

import numpy as np
import torch

from ..distributions.beta import entropy as beta_dist_entropy
from ..distributions.gamma import entropy as gamma_dist_entropy
from ..distributions.multinomial import entropy as multinomial_dist_entropy
from ..utils.constants import DP_DEFAULT_ALPHA_PRIOR_PARAMS, DP_DEFAULT_TRUNCATION_LEVEL
from ..utils.special import digamma
from ..utils.types import TORCH_DTYPE, create_positive_variable, create_random_positive_variable
from .interfaces.trainable import Trainable


def dirichlet_process(num_samples, alpha_prior_params=DP_DEFAULT_ALPHA_PRIOR_PARAMS,
                      truncation_level=DP_DEFAULT_TRUNCATION_LEVEL, mask_size=1, device=None):
    s_1 = float(alpha_prior_params[0])
    s_2 = float(alpha_prior_params[1])

    # same draw order from numpy's global RNG as the reference: logits, gamma_1, gamma_2
    if mask_size == 1:
        logits = torch.tensor(np.random.standard_normal((num_samples, truncation_level)), dtype=TORCH_DTYPE,
                              device=device, requires_grad=True)
        phi_mask = None
    else:
        mask_depth = int(np.divide(num_samples, mask_size))
        mask_indices = np.repeat(np.arange(mask_depth), mask_size)
        phi_mask = torch.nn.functional.one_hot(torch.as_tensor(mask_indices), mask_depth).to(dtype=TORCH_DTYPE, device=device)
        logits = torch.tensor(np.random.standard_normal((mask_depth, truncation_level)), dtype=TORCH_DTYPE,
                              device=device, requires_grad=True)
    gamma_1 = create_random_positive_variable(shape=truncation_level - 1, device=device)
    gamma_2 = create_random_positive_variable(shape=truncation_level - 1, device=device)
    w_1 = create_positive_variable(initial_value=alpha_prior_params[0], device=device)
    w_2 = create_positive_variable(initial_value=alpha_prior_params[1], device=device)

    def phi_value():
        sm = torch.softmax(logits, dim=-1)
        return sm if phi_mask is None else phi_mask @ sm

    def objective_value(phi=None):
        phi = phi_value() if phi is None else phi
        g1, g2, w1, w2 = gamma_1.value, gamma_2.value, w_1.value, w_2.value
        dg12 = digamma(g1 + g2)
        tail = torch.flip(torch.cumsum(torch.flip(phi, [1]), 1), [1]) - phi      # cumsum(exclusive, reverse)
        ev_q_log_p_z_given_v = torch.sum(phi[:, 0:-1] * (digamma(g1) - dg12) + tail[:, 0:-1] * (digamma(g2) - dg12))
        ev_q_log_p_v_given_alpha = (truncation_level - 1.0) * (digamma(w1) - torch.log(w2)) + \
            ((w1 / w2) - 1.0) * torch.sum(digamma(g2) - dg12)
        ev_q_log_p_alpha = s_1 * math.log(s_2) - math.lgamma(s_1) + (s_1 - 1.0) * (digamma(w1) - torch.log(w2)) - \
            s_2 * (w1 / w2)
        entropy_q_z = torch.sum(multinomial_dist_entropy(phi))
        entropy_q_v = torch.sum(beta_dist_entropy(g1, g2))
        entropy_q_alpha = gamma_dist_entropy(w1, w2)
        elbo = ev_q_log_p_z_given_v + ev_q_log_p_v_given_alpha + ev_q_log_p_alpha + entropy_q_z + entropy_q_v + entropy_q_alpha
        return -elbo

    class DirichletProcess(Trainable):
        @property
        def assignments(self):
            return phi_value()

        @property
        def q_z(self):
            return phi_value()

        @property
        def q_v(self):
            return gamma_1.value, gamma_2.value

        @property
        def q_alpha(self):
            return w_1.value, w_2.value

        @property
        def objective(self):
            return objective_value()

        # --- additions over the reference: the trainable leaves and evaluation at a given phi ---
        def objective_at(self, phi):
            return objective_value(phi)

        @property
        def variables(self):
            """(name, leaf tensor) in the reference's tf.Variable creation order."""
            return [("phi_logits", logits), ("gamma1_raw", gamma_1.raw), ("gamma2_raw", gamma_2.raw),
                    ("w1_raw", w_1.raw), ("w2_raw", w_2.raw)]

    return DirichletProcess()
