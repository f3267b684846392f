"""GPU tests at the BASELINE.json shapes: exact parity with the CPU oracle where the oracle finishes in
seconds (a row prefix of the headline configuration C5), and size-independent properties where it does not
(C5 at 131 072 rows, the Frey-faces shape C4 in both modes)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def synth(rng, n, d, q, m, t):
    from oracle.literal import random_params
    y = rng.standard_normal((n, d))
    return y, random_params(rng, n, d, q, m, t)


def test_headline_shape_row_prefix_vs_oracle():
    """C5 (D = 64, Q = 10, M = 128, T = 10) on a 256-row prefix: objective and every gradient block vs oracle/streaming.py."""
    from dp_gp_lvm_b200.models.dp_gp_lvm import PARAM_ORDER, dp_gp_lvm_t
    from oracle import streaming as S
    rng = np.random.default_rng(0)
    n, d, q, m, t = 256, 64, 10, 128, 10
    y, params = synth(rng, n, d, q, m, t)
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, seed=0, device=DEV)
    model.load_variables(params)
    obj, grads = model.value_and_grad()
    ref, gref = S.value_and_grad(y, params, "t", chunk=64)
    assert abs(obj - ref) <= 1e-9 * abs(ref), (obj, ref)
    from conftest import grad_tol
    for k in PARAM_ORDER:
        assert relerr(grads[k].reshape(-1), gref[k].reshape(-1)) < grad_tol(k, 1e-9), k


def test_headline_shape_properties_at_scale():
    """C5 at N = 131 072 (the per-GPU shard of the 8-GPU run): bitwise reproducibility of objective and gradients, and the
    gradient against central differences of the objective along a random direction in every parameter block."""
    from dp_gp_lvm_b200.models.dp_gp_lvm import PARAM_ORDER, dp_gp_lvm_t
    rng = np.random.default_rng(1)
    n, d, q, m, t = 131072, 64, 10, 128, 10
    y, params = synth(rng, n, d, q, m, t)
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, seed=0, device=DEV)
    model.load_variables(params)
    leaves = model.parameters()
    o1 = model.objective; g1 = torch.autograd.grad(o1, leaves)
    o2 = model.objective; g2 = torch.autograd.grad(o2, leaves)
    model.engine.check()
    assert float(o1.item()) == float(o2.item())
    for a, b in zip(g1, g2):
        assert torch.equal(a, b)
    gen = torch.Generator(device=DEV); gen.manual_seed(3)
    for name, p, g in zip(PARAM_ORDER, leaves, g1):
        if name in ("w1_raw", "w2_raw"):
            continue
        direction = torch.randn(p.shape, dtype=torch.float64, device=DEV, generator=gen)
        direction /= direction.norm()
        h = 1e-5 * max(1.0, float(p.detach().norm())) if p.numel() < 1000 else 1e-3
        with torch.no_grad():
            p.add_(h * direction); op = float(model.objective.item())
            p.add_(-2 * h * direction); om = float(model.objective.item())
            p.add_(h * direction)
        fd = (op - om) / (2 * h)
        an = float((g * direction).sum().item())
        assert abs(fd - an) <= 2e-5 * max(abs(an), 1e-3 * abs(float(o1.item())) / max(1.0, h * 1e5)) + 1e-6 * abs(an), (name, fd, an)


def test_frey_shape_both_modes_agree_at_equal_atoms():
    """C4 (N = 1965, D = 560, Q = 10, M = 100, T = 20): too large for the CPU oracle in a test, so it is pinned by the
    identity the reference's own unit test uses (test/unittests/dpgplvm_unitttests.py:547-548): with equal atoms the
    T-mode bound equals the D-mode bound -- objective and the gradients of the shared variables."""
    from dp_gp_lvm_b200.models.dp_gp_lvm import PARAM_ORDER, dp_gp_lvm, dp_gp_lvm_t
    rng = np.random.default_rng(2)
    n, d, q, m, t = 1965, 560, 10, 100, 20
    y, params = synth(rng, n, d, q, m, t)
    for k in ("gamma_atoms_raw", "alpha_atoms_raw", "beta_atoms_raw"):
        params[k] = np.broadcast_to(params[k][:1], params[k].shape).copy()
    np.random.seed(0)
    mt = dp_gp_lvm_t(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, seed=0, device=DEV)
    md = dp_gp_lvm(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, device=DEV)
    mt.load_variables(params); md.load_variables(params)
    ot, gt = mt.value_and_grad()
    od, gd = md.value_and_grad()
    assert abs(ot - od) <= 1e-10 * abs(ot), (ot, od)
    for k in ("x_mean", "x_var_raw", "x_u", "phi_logits", "gamma1_raw", "gamma2_raw"):
        assert relerr(gd[k], gt[k]) < 1e-8, (k, relerr(gd[k], gt[k]))
    # the atom gradients differ by construction (D-mode mixes atoms through phi) but their totals over atoms agree
    for k in ("gamma_atoms_raw", "alpha_atoms_raw", "beta_atoms_raw"):
        assert relerr(gd[k].sum(axis=0), gt[k].sum(axis=0)) < 1e-8, k


@pytest.mark.parametrize("name", ["c4_t", "c4_d64", "c4_d", "c5_d256"])
def test_config_shapes_vs_committed_streaming_oracle(name):
    """BASELINE.json configs[3] (Frey faces shape, test/frey_faces_prediction.py:274-275: N = 1965, D = 560, Q = 10, M = 100,
    T = 20) in T-mode and in D-mode (all 560 kernels, and a 64-column problem), and the headline shape in D-mode on 256
    rows: objective and every gradient block against oracle/streaming.py, whose results are committed fixtures
    (oracle/make_golden_configs.py; the inputs are regenerated from the seed and checked by a checksum)."""
    import os
    from conftest import ROOT, grad_tol, report
    from dp_gp_lvm_b200.models.dp_gp_lvm import PARAM_ORDER, dp_gp_lvm, dp_gp_lvm_t
    from oracle.make_golden_configs import seeded_problem
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated (oracle/make_golden_configs.py %s)" % name)
    z = np.load(path)
    n, d, q, m, t = (int(v) for v in z["shape"])
    mode = str(z["mode"])
    y, params = seeded_problem(int(z["seed"]), n, d, q, m, t)
    assert abs(float(np.abs(y).sum()) - float(z["y_checksum"])) <= 1e-12 * float(z["y_checksum"]), "numpy Generator stream changed"
    np.random.seed(0)
    if mode == "t":
        model = dp_gp_lvm_t(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, seed=0, device=DEV)
    else:
        model = dp_gp_lvm(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, device=DEV)
    model.load_variables(params)
    obj, grads = model.value_and_grad()
    ref = float(z["objective"])
    errs = {k: relerr(grads[k].reshape(-1), z["g_" + k].reshape(-1)) for k in PARAM_ORDER}
    report(name, float("nan"), abs(obj - ref) / abs(ref), errs, 1e-9, 1e-9)
    assert abs(obj - ref) <= 1e-9 * abs(ref), (obj, ref)
    for k in PARAM_ORDER:
        assert errs[k] < grad_tol(k, 1e-9), (k, errs[k])


@pytest.mark.parametrize("bwd_variant", [0, 7])
def test_headline_65536_row_prefix_vs_committed_oracle_result(bwd_variant):
    """(bwd_variant 0: the default DFMA backward; 7: dv / dD on the tcgen05 tensor cores.)  SURVEY.md 8d: the headline configuration on the first 65 536 rows of bench.py's synthetic problem against the CPU
    streaming oracle.  The oracle needs ~1 h for this on 8 cores, so its result is a committed fixture
    (tests/golden/c5_prefix65536.npz, written by oracle/make_c5_golden.py): objective, every small gradient block in
    full, and the per-row blocks as column sums, norm and every 257th row."""
    import os
    import bench
    from conftest import ROOT, grad_tol
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm_t
    path = os.path.join(ROOT, "tests", "golden", "c5_prefix65536.npz")
    if not os.path.exists(path):
        pytest.skip("fixture not generated (oracle/make_c5_golden.py)")
    z = np.load(path)
    n, stride = int(z["n"]), int(z["stride"])
    shape = bench.SHAPE
    y, params = bench.synthetic(n, 0, shape)
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=shape["q"], num_inducing_points=shape["m"], truncation_level=shape["t"],
                        seed=0, device=DEV, bwd_variant=bwd_variant)
    model.load_variables(params)
    obj, grads = model.value_and_grad()
    ref = float(z["objective"])
    assert abs(obj - ref) <= 1e-9 * abs(ref), (obj, ref)
    for k, g in grads.items():
        g = np.asarray(g)
        if k in ("x_mean", "x_var_raw"):
            assert relerr(g[::stride], z["rows_" + k]) < 1e-9, k
            assert relerr(g.sum(axis=0), z["sum_" + k]) < 1e-9, k
            assert abs(np.sqrt((g ** 2).sum()) - float(z["norm_" + k])) <= 1e-9 * float(z["norm_" + k]), k
        else:
            assert relerr(g.reshape(-1), z["grad_" + k].reshape(-1)) < grad_tol(k, 1e-9), (k, relerr(g.reshape(-1), z["grad_" + k].reshape(-1)))
