"""B200-native DP-GP-LVM bound (ELBO + gradients): hand-written sm_100a CUDA kernels behind the C ABI of
include/dpgp.h, with the reference's Python model/kernel API (AndrewRLawrence/dp_gp_lvm) on top."""
from ._lib import MODE_D, MODE_T, DpgpError, NotPositiveDefiniteError  # noqa: F401
