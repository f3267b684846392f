"""CPU tests: the oracle against the fixtures produced by the reference's own code (tests/golden/,
written by oracle/make_golden.py) and against the reference's known-answer ("naive") definitions.
These pin the oracle; the GPU tests then compare the CUDA path with the oracle."""
import numpy as np
import pytest
import torch

from conftest import MODEL_CASES, golden_params, kuu_condition, load_golden, tolerances
from oracle import literal as L
from oracle import naive
from oracle import streaming as S

REL_OBJ = 1e-12      # oracle vs reference-over-shim: same op sequence, expect rounding-level agreement
REL_GRAD = 1e-10


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("name", ["kernel_b1", "kernel_b7"])
def test_kernel_literal_vs_reference(name):
    z = load_golden(name)
    t = lambda k: torch.as_tensor(z[k])
    assert relerr(L.k_uu(t("x_u"), t("gamma"), t("alpha")).numpy(), z["k_uu"]) < 1e-13
    assert relerr(L.psi_0(z["x_mean"].shape[0], t("alpha")).numpy(), z["psi_0"]) < 1e-15
    assert relerr(L.psi_1(t("x_u"), t("x_mean"), t("x_var"), t("gamma"), t("alpha")).numpy(), z["psi_1"]) < 1e-13
    assert relerr(L.psi_2(t("x_u"), t("x_mean"), t("x_var"), t("gamma"), t("alpha")).numpy(), z["psi_2"]) < 1e-13


def test_kernel_naive_vs_reference():
    """the reference's own known-answer definitions (kernel_unittests.py:14-147), B = 1 and one of B = 7"""
    for name, b in (("kernel_b1", 0), ("kernel_b7", 3)):
        z = load_golden(name)
        g, a, be = z["gamma"][b], float(z["alpha"][b, 0]), float(z["beta"][b, 0])
        n = 40   # rows used for the O(N M^2 Q) python loops
        assert relerr(naive.covariance_matrix(z["x_u"], g, a, be, include_jitter=True), z["k_uu"][b]) < 1e-12
        assert relerr(naive.covariance_matrix(z["x0"], g, a, be, include_noise=True, include_jitter=True), z["k_xx"][b]) < 1e-12
        assert relerr(naive.covariance_matrix(z["x0"], g, a, be, x1=z["x1"], include_noise=True, include_jitter=True), z["k_xz"][b]) < 1e-12
        assert relerr(naive.psi_1(z["x_mean"], z["x_var"], z["x_u"], g, a), z["psi_1"][b]) < 1e-12
        assert relerr(naive.psi_0(z["x_mean"].shape[0], a), z["psi_0"][b]) < 1e-15
        if b == 0:
            assert relerr(naive.psi_2(z["x_mean"], z["x_var"], z["x_u"], g, a), z["psi_2"][b]) < 1e-12
        else:
            # partial-N check against the literal oracle on the same rows
            ref = L.psi_2(torch.as_tensor(z["x_u"]), torch.as_tensor(z["x_mean"][:n]), torch.as_tensor(z["x_var"][:n]),
                          torch.as_tensor(z["gamma"][b:b + 1]), torch.as_tensor(z["alpha"][b:b + 1])).numpy()[0]
            assert relerr(naive.psi_2(z["x_mean"][:n], z["x_var"][:n], z["x_u"], g, a), ref) < 1e-12


@pytest.mark.parametrize("name", ["dp_n10_t20", "dp_n12_t5_mask3"])
def test_dp_vs_reference(name):
    z = load_golden(name)
    d, mask = int(z["num_dims"]), int(z["mask_size"])
    leaves = {k: torch.tensor(z[k], requires_grad=True) for k in ("phi_logits", "gamma1_raw", "gamma2_raw", "w1_raw", "w2_raw")}
    phi = L.phi_from_logits(leaves["phi_logits"], d, mask)
    obj = L.dp_objective(phi, L.softplus(leaves["gamma1_raw"]), L.softplus(leaves["gamma2_raw"]),
                         L.softplus(leaves["w1_raw"]), L.softplus(leaves["w2_raw"]),
                         float(z["alpha_prior"][0]), float(z["alpha_prior"][1]))
    assert abs(float(obj.detach()) - float(z["objective"])) <= REL_OBJ * abs(float(z["objective"]))
    grads = torch.autograd.grad(obj, list(leaves.values()))
    for k, g in zip(leaves, grads):
        assert relerr(g.numpy(), z["grad_" + k]) < REL_GRAD, k
    # the reference's known-answer definition (dp_unittests.py:13-131)
    sp = lambda x: np.log1p(np.exp(x))
    elbo = naive.dp_elbo(phi.detach().numpy(), sp(z["gamma1_raw"]), sp(z["gamma2_raw"]), float(sp(z["w1_raw"])),
                         float(sp(z["w2_raw"])), float(z["alpha_prior"][0]), float(z["alpha_prior"][1]))
    assert abs(-elbo - float(z["objective"])) <= 1e-10 * abs(float(z["objective"]))


@pytest.mark.parametrize("case", MODEL_CASES)
@pytest.mark.parametrize("mode", ["t", "d"])
def test_literal_and_streaming_vs_reference(mode, case):
    z = load_golden("%s_%s" % (mode, case))
    params = golden_params(z)
    mask = int(z["mask_size"])
    fn = L.objective_t if mode == "t" else L.objective_d
    obj, grads = L.value_and_grad(fn, z["y"], params, z["alpha_prior"], mask)
    ref = float(z["objective"])
    assert abs(obj - ref) <= REL_OBJ * abs(ref)
    for k in L.PARAM_ORDER:
        if z["g_" + k].size:
            assert relerr(grads[k], z["g_" + k]) < REL_GRAD, k
    # the streaming form re-associates the M x M chain: conditioning floor applies (c1: kappa ~ 1e9)
    tol_obj, tol_grad = tolerances(kuu_condition(z), base_obj=REL_OBJ)
    obj_s, grads_s = S.value_and_grad(z["y"], params, mode, z["alpha_prior"], mask, chunk=17)
    assert abs(obj_s - ref) <= tol_obj * abs(ref)
    for k in L.PARAM_ORDER:
        if z["g_" + k].size:
            assert relerr(grads_s[k], z["g_" + k]) < tol_grad, k


def test_dmode_naive_composition():
    """objective = -(dp_elbo + sum_d F_naive(y_d; mixed hyper-parameters) - KL + hyper-prior),
    test/unittests/dpgplvm_unitttests.py:78-126 (rtol 1e-7 there)."""
    z = load_golden("d_d2t1")
    p = golden_params(z)
    sp = lambda x: np.log1p(np.exp(x))
    phi = z["assignments"]
    val = naive.dmode_objective(z["y"], p["x_mean"], sp(p["x_var_raw"]), p["x_u"], phi, sp(p["gamma1_raw"]),
                                sp(p["gamma2_raw"]), float(sp(p["w1_raw"])), float(sp(p["w2_raw"])),
                                float(z["alpha_prior"][0]), float(z["alpha_prior"][1]),
                                sp(p["gamma_atoms_raw"]), sp(p["alpha_atoms_raw"]), sp(p["beta_atoms_raw"]))
    assert abs(val - float(z["objective"])) <= 1e-7 * abs(float(z["objective"]))
    z = load_golden("d_mask3")
    p = golden_params(z)
    val = naive.dmode_objective(z["y"], p["x_mean"], sp(p["x_var_raw"]), p["x_u"], z["assignments"], sp(p["gamma1_raw"]),
                                sp(p["gamma2_raw"]), float(sp(p["w1_raw"])), float(sp(p["w2_raw"])),
                                float(z["alpha_prior"][0]), float(z["alpha_prior"][1]),
                                sp(p["gamma_atoms_raw"]), sp(p["alpha_atoms_raw"]), sp(p["beta_atoms_raw"]))
    assert abs(val - float(z["objective"])) <= 1e-7 * abs(float(z["objective"]))


def test_t_equals_d_at_equal_atoms():
    """dpgplvm_unitttests.py:547-548: the two formulations coincide at the reference's initialisation."""
    a, b = load_golden("t_init"), load_golden("d_init")
    assert abs(float(a["objective"]) - float(b["objective"])) < 1e-10 * abs(float(a["objective"]))


def test_oracle_uses_an_accurate_trigamma():
    """torch differentiates digamma with torch.special.polygamma(1, x), whose float64 series is cut after the x^-7 term
    (relative error up to ~5e-10; digamma itself is accurate to 1e-15).  The cancellation in d ELBO / d w_1 amplifies
    that to as much as 4e-7 at the BASELINE shapes, so the oracle's digamma (oracle/special.py) back-propagates through
    scipy's polygamma instead -- checked here, together with the reason it is needed."""
    import torch
    from scipy.special import digamma, polygamma
    from oracle.special import digamma as oracle_digamma
    x = np.array([0.3, 1.0, 1.3132616875182228, 2.5, 5.0, 7.7])
    t = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    y = oracle_digamma(t)
    (g,) = torch.autograd.grad(y.sum(), t)
    assert np.abs(g.numpy() - polygamma(1, x)).max() <= 1e-15 * polygamma(1, x).max()
    assert np.abs(y.detach().numpy() - digamma(x)).max() < 1e-14
    tri = torch.special.polygamma(1, t.detach()).numpy()
    err = np.abs(tri - polygamma(1, x)) / polygamma(1, x)
    assert err.max() > 1e-11, "torch's trigamma got better: oracle/special.py is no longer needed"


def test_device_special_functions_on_the_host(tmp_path):
    """csrc/special.cuh (digamma / trigamma / softplus / sigmoid used by the fused small-variable kernels) compiled for the
    host with g++ and pinned against scipy: the product's closed-form DP gradients rest on these (torch's trigamma, which
    the oracle differentiates with, is only good to ~5e-10)."""
    import ctypes as C
    import os
    import shutil
    import subprocess
    from scipy.special import digamma, expit, polygamma
    from conftest import ROOT
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = tmp_path / "special_host.cpp"
    src.write_text('#define DPGP_HD\n#include <cmath>\nusing namespace std;\n#include "special.cuh"\n'
                   'extern "C" {\n'
                   'double h_digamma(double x) { return dpgp::digamma_pos(x); }\n'
                   'double h_trigamma(double x) { return dpgp::trigamma_pos(x); }\n'
                   'double h_softplus(double x) { return dpgp::softplus_d(x); }\n'
                   'double h_sigmoid(double x) { return dpgp::sigmoid_d(x); }\n}\n')
    so = tmp_path / "special_host.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "dp_gp_lvm_b200", "csrc"),
                    str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    for f in ("h_digamma", "h_trigamma", "h_softplus", "h_sigmoid"):
        getattr(lib, f).restype = C.c_double; getattr(lib, f).argtypes = [C.c_double]
    xs = np.concatenate([np.logspace(-6, 3, 400), np.linspace(0.01, 30.0, 997), [1.0, 1.4616321449683623, 2.0, 10.0]])
    dg = np.array([lib.h_digamma(float(x)) for x in xs]); tg = np.array([lib.h_trigamma(float(x)) for x in xs])
    ref_d, ref_t = digamma(xs), polygamma(1, xs)
    assert (np.abs(dg - ref_d) / np.maximum(1.0, np.abs(ref_d))).max() < 4e-15        # absolute near the root at 1.4616
    assert (np.abs(tg - ref_t) / ref_t).max() < 4e-15
    rs = np.linspace(-40.0, 40.0, 801)
    sp = np.array([lib.h_softplus(float(r)) for r in rs]); sg = np.array([lib.h_sigmoid(float(r)) for r in rs])
    assert (np.abs(sp - np.logaddexp(0.0, rs)) / np.logaddexp(0.0, rs)).max() < 4e-16 * 4
    assert (np.abs(sg - expit(rs)) / expit(rs)).max() < 1e-15
