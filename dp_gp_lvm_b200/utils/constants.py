"""Constants of the reference that the bound depends on (src/utils/constants.py:86-121), verbatim values.
The reference's path/plot/GUI constants (and its sys.path lookup, constants.py:76) are out of scope."""
import numpy as np

NP_DTYPE = np.float64

OPT_DEFAULT_LEARNING_RATE = 0.05
OPT_DEFAULT_ITERS = 901
OPT_MAX_ITERS = 2501

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
