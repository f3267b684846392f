"""Constants of the reference that the bound depends on (src/utils/constants.py:86-121), verbatim values.
The reference's path/plot/GUI constants (and its sys.path lookup, constants.py:76) are out of scope."""
from enum import Enum

import numpy as np

NP_DTYPE = np.float64


class ResultKeys(Enum):
    """Array names in result .npz files (reference src/utils/constants.py:38-73), so that the reference's
    analyse_*.py readers consume files written by train.save_results unchanged."""
    ORIGINAL_DATA = 'original_data'
    RANDOMISED_DATA = 'randomised_data'
    NORMALISED_DATA = 'normalised_data'
    TRAINING_DATA = 'y_train'
    TRAINING_INPUT_MEAN = 'x_mean'
    TRAINING_INPUT_COVAR = 'x_covar'
    INDUCING_INPUT = 'x_u'
    TEST_DATA = 'y_test'
    TEST_INPUT_MEAN = 'x_test_mean'
    TEST_INPUT_COVAR = 'x_test_covar'
    ARD_WEIGHTS = 'ard_weights'
    SIGNAL_VARIANCE = 'signal_variance'
    NOISE_PRECISION = 'noise_precision'
    DP_ASSIGNMENTS = 'assignments'
    Q_ALPHA_W1 = 'q_alpha_w1'
    Q_ALPHA_W2 = 'q_alpha_w2'
    Q_V_A = 'q_v_a'
    Q_V_B = 'q_v_b'
    ARD_WEIGHTS_ATOMS = 'gamma_atoms'
    SIGNAL_VARIANCE_ATOMS = 'alpha_atoms'
    NOISE_PRECISION_ATOMS = 'beta_atoms'


OPT_DEFAULT_LEARNING_RATE = 0.05
OPT_DEFAULT_ITERS = 901
OPT_MAX_ITERS = 2501

MAX_MC_SAMPLES = 5000

GP_DEFAULT_JITTER = 1.0e-8
GP_INIT_GAMMA = 1.0
GP_INIT_ALPHA = 1.0
GP_INIT_BETA = 1.0

GP_LVM_DEFAULT_LATENT_DIMENSIONS = 10
GP_LVM_DEFAULT_NUM_INDUCING_POINTS = 25
GP_LVM_MAX_LATENT_DIMENSIONS = 25

DP_MAX_NUM_SAMPLES = 5000
DP_MAX_TRUNCATION_LEVEL = 50
DP_DEFAULT_ALPHA = 1.0
DP_DEFAULT_ALPHA_PRIOR_PARAMS = np.array([1.0, 1.0], dtype=NP_DTYPE)
DP_DEFAULT_TRUNCATION_LEVEL = 8

NP_SEED = 1
