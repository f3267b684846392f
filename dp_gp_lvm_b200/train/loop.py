"""Training loop and result files of the reference's experiment scripts.

`train` is the loop of test/synthetic_data_hard_test.py:147-160 (Adam, print the objective every 100 iterations,
return the wall-clock optimisation time); `save_results` writes the .npz with the keys those scripts and the
reference's analyse_*.py readers use (src/utils/constants.py:38-73, test/synthetic_data_hard_test.py:163-172)."""
from time import time

import numpy as np
import torch

from ..utils.constants import OPT_DEFAULT_ITERS, OPT_DEFAULT_LEARNING_RATE, ResultKeys
from .optimiser import AdamOptimizer


def train(model, learning_rate=OPT_DEFAULT_LEARNING_RATE, train_iter=OPT_DEFAULT_ITERS, print_every=100, name='GP-DP',
          use_cuda_graph=False, verbose=True, callback=None):
    """Returns (train_opt_time in seconds, list of (iteration, objective) pairs that were printed)."""
    train_op = AdamOptimizer(learning_rate=learning_rate, use_cuda_graph=use_cuda_graph).minimize(loss=model)
    history = []
    torch.cuda.synchronize()
    start_time = time()
    if verbose:
        print('\nTraining {}..'.format(name))
    for c in range(train_iter):
        train_op.run()
        if print_every and (c % print_every) == 0:
            # the reference prints the objective AFTER the update of iteration c (a fresh session.run)
            with torch.no_grad():                      # value only: skips the statistics backward
                val = float(model.objective.item())
            model.engine.check()                       # the reference's tf.cholesky would have aborted by now
            history.append((c, val))
            if verbose:
                print('  {} opt iter {:5}: {}'.format(name, c, val))
            if callback is not None:
                callback(c, val)
    model.engine.check()
    torch.cuda.synchronize()
    train_opt_time = time() - start_time
    with torch.no_grad():
        final = float(model.objective.item())
    history.append((train_iter - 1, final))
    if verbose:
        print('Final iter {:5}:'.format(train_iter - 1))
        print('  {}: {}'.format(name, final))
        print('Time to optimise: {} s'.format(train_opt_time))
    return train_opt_time, history


def _np(t):
    return t.detach().cpu().numpy()


def save_results(model, file_name, y_train, train_opt_time=None, **extra):
    """np.savez with the reference's keys (ResultKeys); DP-GP-LVM models also store the atoms and the assignments."""
    x_mean, x_covar = model.q_x
    out = {ResultKeys.TRAINING_DATA.value: np.asarray(y_train),
           ResultKeys.ARD_WEIGHTS.value: _np(model.ard_weights),
           ResultKeys.NOISE_PRECISION.value: _np(model.noise_precision),
           ResultKeys.SIGNAL_VARIANCE.value: _np(model.signal_variance),
           ResultKeys.INDUCING_INPUT.value: _np(model.inducing_input),
           ResultKeys.TRAINING_INPUT_MEAN.value: _np(x_mean),
           ResultKeys.TRAINING_INPUT_COVAR.value: _np(x_covar)}
    if hasattr(model, 'assignments'):
        g, a, b = model.dp_atoms
        out.update({ResultKeys.DP_ASSIGNMENTS.value: _np(model.assignments),
                    ResultKeys.ARD_WEIGHTS_ATOMS.value: _np(g), ResultKeys.SIGNAL_VARIANCE_ATOMS.value: _np(a),
                    ResultKeys.NOISE_PRECISION_ATOMS.value: _np(b)})
    if train_opt_time is not None:
        out['train_opt_time'] = train_opt_time
    out.update(extra)
    np.savez(file_name, **out)
    return out
