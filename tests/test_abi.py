"""CPU tests of the drop-in boundary: the C-ABI library loads and exports exactly what include/dpgp.h
declares (no compute calls -- there is no GPU here), and the host API mirrors the reference's names."""
import os
import re

import pytest

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "dpgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dp_gp_lvm_b200 import _lib
    lib = _lib.lib()
    declared = header_functions()
    assert declared, "no functions parsed from include/dpgp.h"
    for name in declared:
        assert hasattr(lib, name), "libdpgp.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared


@pytest.mark.parametrize("nb", [1, 2, 3, 4, 7, 13, 16, 17, 25, 32])
def test_fused_backward_schedule_is_a_conflict_free_cover(nb):
    """Host-only: the round schedule of psi2_bwd_fused_kernel covers every 8x8 block bi <= bj exactly once and
    no two blocks of a round share an m-block (the property its shared, barrier-ordered d r accumulation needs)."""
    import ctypes as C
    from dp_gp_lvm_b200 import _lib
    lib = _lib.lib()
    buf = (C.c_ushort * 8192)()
    nr = lib.dpgp_fused_schedule(nb, buf, 8192)
    assert nr > 0
    seen = set()
    for r in range(nr):
        used = set()
        for w in range(8):
            it = buf[r * 8 + w]
            if it == 0xffff:
                continue
            bi, bj = it >> 8, it & 255
            assert bi <= bj < nb and (bi, bj) not in seen
            seen.add((bi, bj))
            assert bi not in used and bj not in used
            used.update((bi, bj))
    assert len(seen) == nb * (nb + 1) // 2
    assert lib.dpgp_fused_schedule(0, None, 0) < 0 and lib.dpgp_fused_schedule(33, None, 0) < 0


def test_reference_api_surface():
    """Names a user of the reference imports."""
    from dp_gp_lvm_b200.kernels.interfaces.kernel import AbstractKernel, Kernel, KernelHyperparameters
    from dp_gp_lvm_b200.kernels.rbf_kernel import k_ard_rbf, k_mahalanobis_rbf, k_rbf
    from dp_gp_lvm_b200.models.dirichlet_process import dirichlet_process
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    from dp_gp_lvm_b200.models.interfaces.trainable import Trainable
    import inspect
    assert [m.value for m in KernelHyperparameters][:3] == ["gamma", "alpha", "beta"]
    assert list(inspect.signature(dp_gp_lvm).parameters)[:6] == [
        "y_train", "num_latent_dims", "num_inducing_points", "truncation_level", "alpha_prior_params", "mask_size"]
    assert list(inspect.signature(dp_gp_lvm_t).parameters)[:7] == [
        "y_train", "num_latent_dims", "num_inducing_points", "truncation_level", "alpha_prior_params", "mask_size", "seed"]
    sig = inspect.signature(dp_gp_lvm)
    assert sig.parameters["num_latent_dims"].default == 10 and sig.parameters["num_inducing_points"].default == 25
    assert sig.parameters["truncation_level"].default == 8 and sig.parameters["mask_size"].default == 1
    assert list(inspect.signature(dirichlet_process).parameters)[:4] == [
        "num_samples", "alpha_prior_params", "truncation_level", "mask_size"]
    for fn in (k_rbf, k_mahalanobis_rbf):
        with pytest.raises(NotImplementedError):
            fn(*([None] * len(inspect.signature(fn).parameters)))
    assert issubclass(Kernel, AbstractKernel) and hasattr(Trainable, "objective")
    assert callable(k_ard_rbf)


def test_next_rows_api_surface():
    """the training-loop, prediction and BGP-LVM entry points keep the reference's names and arguments."""
    import inspect
    from dp_gp_lvm_b200.models.gaussian_process import bayesian_gp_lvm
    from dp_gp_lvm_b200.train import AdamOptimizer, save_results, train
    from dp_gp_lvm_b200.utils.constants import ResultKeys
    assert list(inspect.signature(bayesian_gp_lvm).parameters)[:5] == [
        "y_train", "kernel", "num_latent_dims", "num_inducing_points", "num_latent_samples"]
    sig = inspect.signature(AdamOptimizer.__init__)
    assert [sig.parameters[k].default for k in ("learning_rate", "beta1", "beta2", "epsilon")] == [0.001, 0.9, 0.999, 1e-08]
    assert callable(train) and callable(save_results)
    assert ResultKeys.TRAINING_INPUT_MEAN.value == "x_mean" and ResultKeys.ARD_WEIGHTS_ATOMS.value == "gamma_atoms"
    import numpy as np
    with pytest.raises(AssertionError):
        bayesian_gp_lvm(y_train=np.zeros((10, 4)), num_latent_dims=4, num_inducing_points=3)


def test_argument_validation_is_assertion_error():
    """The reference validates with Python asserts (dp_gp_lvm.py:52-59, :544-559) before touching the device."""
    import numpy as np
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    y = np.zeros((20, 6))
    with pytest.raises(AssertionError):
        dp_gp_lvm_t(y_train=y, num_latent_dims=6, num_inducing_points=5, truncation_level=3)     # Q < D required
    with pytest.raises(AssertionError):
        dp_gp_lvm_t(y_train=y.tolist(), num_latent_dims=2, num_inducing_points=5, truncation_level=3)
    with pytest.raises(AssertionError):
        dp_gp_lvm(y_train=y, num_latent_dims=2, num_inducing_points=21, truncation_level=3)      # M <= N
    with pytest.raises(AssertionError):
        dp_gp_lvm(y_train=y, num_latent_dims=2, num_inducing_points=5, truncation_level=7)       # T <= min(N, D)
    with pytest.raises(AssertionError):
        dp_gp_lvm_t(y_train=y, num_latent_dims=2, num_inducing_points=5, truncation_level=3, seed=-1)


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not fall back to the oracle."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dp_gp_lvm_b200.engine import BoundEngine
    with pytest.raises(RuntimeError):
        BoundEngine(10, 4, 2, 5, 3, 0)
    import dp_gp_lvm_b200
    pkg = os.path.dirname(dp_gp_lvm_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), "%s imports the oracle" % f
