"""TEST INFRASTRUCTURE ONLY -- placeholder so `import tensorflow_probability as tfp` in the reference
resolves.  The DP-GP-LVM bound never calls into it (only the commented-out SVI model and the
Monte-Carlo psi option of bayesian_gp_lvm do), so every attribute access fails loudly."""


class _Missing:
    def __getattr__(self, name):
        raise NotImplementedError("tensorflow_probability.%s is not provided by the oracle shim" % name)


distributions = _Missing()
