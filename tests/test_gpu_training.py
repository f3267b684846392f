"""GPU tests of the caller of the hot path: the Adam step (dpgp_adam), the training loop that
mirrors the reference scripts, CUDA-graph replay of a whole iteration, and the .npz result format.

The checker for the trajectory is the CPU oracle (oracle/literal.py, the reference's op sequence under autograd)
driven by a numpy restatement of TensorFlow-1's Adam (tf.train.AdamOptimizer: lr_t = lr sqrt(1-b2^t)/(1-b1^t),
theta -= lr_t m / (sqrt(v) + eps))."""
import numpy as np
import pytest
import torch

from conftest import golden_params, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def tf_adam_numpy(theta, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m[...] = b1 * m + (1.0 - b1) * g
    v[...] = b2 * v + (1.0 - b2) * g * g
    theta[...] = theta - lr_t * m / (np.sqrt(v) + eps)


def build(name, mode="t", **kw):
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    z = load_golden(name)
    p = golden_params(z)
    q = p["x_mean"].shape[1]; m = p["x_u"].shape[0]; t = p["gamma_atoms_raw"].shape[0]
    np.random.seed(0)
    fac = dp_gp_lvm_t if mode == "t" else dp_gp_lvm
    extra = dict(seed=0) if mode == "t" else {}
    model = fac(y_train=z["y"], num_latent_dims=q, num_inducing_points=m, truncation_level=t,
                alpha_prior_params=z["alpha_prior"], mask_size=int(z["mask_size"]), device=DEV, **extra, **kw)
    model.load_variables(p)
    return model, z, p


@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
def test_adam_kernel_is_tensorflow_adam(n):
    from dp_gp_lvm_b200.engine import BoundEngine, MODE_T
    eng = BoundEngine(8, 3, 2, 4, 2, MODE_T, device=DEV)
    rng = np.random.default_rng(n)
    theta = rng.standard_normal(n); m = np.zeros(n); v = np.zeros(n)
    th_d = torch.tensor(theta, device=DEV); m_d = torch.zeros(n, dtype=torch.float64, device=DEV); v_d = torch.zeros_like(m_d)
    step = torch.zeros((), dtype=torch.int64, device=DEV)
    for t in range(1, 6):
        g = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 3, size=n)
        tf_adam_numpy(theta, g, m, v, t, 0.05)
        step += 1
        eng.adam(th_d, torch.tensor(g, device=DEV), m_d, v_d, step, 0.05)
    assert np.abs(th_d.cpu().numpy() - theta).max() <= 1e-14 * max(1.0, np.abs(theta).max())
    assert np.abs(m_d.cpu().numpy() - m).max() <= 1e-15 * max(1.0, np.abs(m).max())
    assert np.abs(v_d.cpu().numpy() - v).max() <= 1e-15 * max(1.0, np.abs(v).max())


def test_adam_multi_is_adam_per_tensor():
    """dpgp_adam_multi (one launch for all tensors) == dpgp_adam tensor by tensor, bitwise."""
    from dp_gp_lvm_b200.engine import BoundEngine, MODE_T
    eng = BoundEngine(8, 3, 2, 4, 2, MODE_T, device=DEV)
    gen = torch.Generator(device=DEV); gen.manual_seed(5)
    shapes = [(1,), (), (7, 3), (4096,), (100003,), (2, 5)]
    mk = lambda: [torch.randn(s, dtype=torch.float64, device=DEV, generator=gen) for s in shapes]
    th_a = mk(); th_b = [t.clone() for t in th_a]
    m_a = [torch.zeros_like(t) for t in th_a]; v_a = [torch.zeros_like(t) for t in th_a]
    m_b = [torch.zeros_like(t) for t in th_a]; v_b = [torch.zeros_like(t) for t in th_a]
    step = torch.zeros((), dtype=torch.int64, device=DEV)
    for _ in range(4):
        g = mk()
        step += 1
        eng.adam_multi(th_a, g, m_a, v_a, step, 0.05)
        for t, gg, m, v in zip(th_b, g, m_b, v_b):
            eng.adam(t, gg, m, v, step, 0.05)
    for a, b in zip(th_a + m_a + v_a, th_b + m_b + v_b):
        assert torch.equal(a, b)


@pytest.mark.parametrize("mode,case", [("t", "q10"), ("d", "c3s")])
def test_training_trajectory_vs_oracle(mode, case):
    """15 Adam iterations on the GPU path vs 15 iterations of the CPU oracle's gradients + numpy TF-Adam."""
    from oracle import literal as L
    from dp_gp_lvm_b200.models.dp_gp_lvm import PARAM_ORDER
    from dp_gp_lvm_b200.train import AdamOptimizer
    model, z, p0 = build("%s_%s" % (mode, case), mode)
    lr, iters = 0.01, 15
    train_op = AdamOptimizer(learning_rate=lr).minimize(loss=model)
    gpu_traj = []
    for _ in range(iters):
        train_op.run()
        gpu_traj.append(float(train_op.objective.item()))
    model.engine.check()
    # oracle trajectory
    fn = L.objective_t if mode == "t" else L.objective_d
    params = {k: np.array(v, dtype=np.float64, copy=True) for k, v in p0.items()}
    ms = {k: np.zeros_like(v) for k, v in params.items()}; vs = {k: np.zeros_like(v) for k, v in params.items()}
    ref_traj = []
    for t in range(1, iters + 1):
        obj, grads = L.value_and_grad(fn, z["y"], params, tuple(z["alpha_prior"]), int(z["mask_size"]))
        ref_traj.append(obj)
        for k in PARAM_ORDER:
            tf_adam_numpy(params[k], grads[k].reshape(params[k].shape), ms[k], vs[k], t, lr)
    gpu_traj, ref_traj = np.array(gpu_traj), np.array(ref_traj)
    assert ref_traj[-1] < ref_traj[0], "the oracle's objective must decrease over the run"
    assert np.abs(gpu_traj - ref_traj).max() <= 1e-8 * np.abs(ref_traj).max(), (gpu_traj, ref_traj)
    final = model.variables
    for k in PARAM_ORDER:
        a = final[k].detach().cpu().numpy().reshape(params[k].shape)
        assert np.abs(a - params[k]).max() <= 1e-7 * max(1.0, np.abs(params[k]).max()), k


def test_cuda_graph_replay_is_the_same_iteration():
    """A captured-and-replayed iteration gives bitwise the trajectory of eager iterations (all kernels deterministic)."""
    from dp_gp_lvm_b200.train import AdamOptimizer
    m1, _, _ = build("t_q10")
    m2, _, _ = build("t_q10")
    op1 = AdamOptimizer(learning_rate=0.02).minimize(loss=m1)
    op2 = AdamOptimizer(learning_rate=0.02, use_cuda_graph=True).minimize(loss=m2)
    for _ in range(8):
        op1.run(); op2.run()
    assert op1.iterations == op2.iterations == 8
    assert float(op1.objective.item()) == float(op2.objective.item())
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(a, b)


def test_train_loop_and_result_file(tmp_path):
    """train() mirrors test/synthetic_data_hard_test.py:147-160; save_results writes the reference's .npz keys."""
    from dp_gp_lvm_b200.train import save_results, train
    model, z, _ = build("t_c3s")
    t_opt, hist = train(model, learning_rate=0.01, train_iter=21, print_every=10, verbose=False, use_cuda_graph=True)
    assert [c for c, _ in hist] == [0, 10, 20, 20] and hist[-1][1] < hist[0][1] and t_opt > 0
    f = tmp_path / "gpdp_test.npz"
    save_results(model, str(f), z["y"], train_opt_time=t_opt)
    r = np.load(str(f))
    n, d = z["y"].shape
    q = z["p_x_mean"].shape[1]; m = z["p_x_u"].shape[0]; t = z["p_gamma_atoms_raw"].shape[0]
    for key in ("y_train", "ard_weights", "noise_precision", "signal_variance", "x_u", "x_mean", "x_covar", "assignments",
                "gamma_atoms", "alpha_atoms", "beta_atoms", "train_opt_time"):
        assert key in r.files, key
    assert r["x_mean"].shape == (n, q) and r["x_covar"].shape == (n, q, q) and r["x_u"].shape == (m, q)
    assert r["assignments"].shape == (d, t) and r["gamma_atoms"].shape == (t, q) and r["ard_weights"].shape == (t, q)


def test_engine_released_during_a_capture_does_not_invalidate_it():
    """An engine that dies (Python's cycle collector, `del`) while a CUDA-graph capture is under way must not free device
    memory or synchronise a stream inside the capture: its handle is parked and destroyed by the next create / close outside
    one (dp_gp_lvm_b200/engine.py).  bench.py hit this as cudaErrorStreamCaptureInvalidated when the collector ran inside the
    capture of the next model's iteration."""
    from dp_gp_lvm_b200 import engine as E
    doomed = E.BoundEngine(64, 4, 2, 8, 3, E.MODE_T, device=DEV)
    x = torch.zeros(8, dtype=torch.float64, device=DEV)
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        x += 1.0
        doomed.close()
        assert len(E._DEFERRED_DESTROY) == 1
        x += 1.0
    g.replay()
    torch.cuda.synchronize()
    assert float(x[0]) == 2.0
    other = E.BoundEngine(64, 4, 2, 8, 3, E.MODE_T, device=DEV)          # drains the parked handle
    assert E._DEFERRED_DESTROY == []
    other.close()
    # and the training op collects before it captures and keeps the collector off inside
    model, z, p = build("t_q10")
    trash = [build("t_q10")[0] for _ in range(2)]
    for t_ in trash:
        t_.cycle = t_                                                     # only the cycle collector can free these
    del trash, t_
    from dp_gp_lvm_b200.train import AdamOptimizer
    op = AdamOptimizer(learning_rate=0.01, use_cuda_graph=True).minimize(loss=model)
    for _ in range(3):
        op.run()
    torch.cuda.synchronize()
    model.engine.check()
    assert np.isfinite(float(op.objective.item()))


@pytest.mark.parametrize("mode,name", [("t", "t_q10"), ("d", "d_c3s"), ("t", "t_t1")])
def test_fused_adam_iteration_equals_the_autograd_path(mode, name):
    """TrainOp's default path (models/dp_gp_lvm.py: fused_adam_iteration -> dpgp_train_tail: no autograd, no torch glue) against
    the autograd path of the same optimiser (forced by passing an objective_fn): same objective trace, same variables."""
    from dp_gp_lvm_b200.train import AdamOptimizer
    m1, _, _ = build(name, mode)
    m2, _, _ = build(name, mode)
    op1 = AdamOptimizer(learning_rate=0.02).minimize(loss=m1)
    op2 = AdamOptimizer(learning_rate=0.02).minimize(loss=m2, objective_fn=lambda: m2.objective)
    assert op1._fast and not op2._fast
    for it in range(6):
        op1.run(); op2.run()
        a, b = float(op1.objective.item()), float(op2.objective.item())
        assert abs(a - b) <= 1e-11 * abs(b), (it, a, b)
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        if p1.numel():
            den = max(float(p2.abs().max()), 1e-300)
            assert float((p1 - p2).abs().max()) / den < 1e-11
    assert int(op1.step.item()) == 6 == int(op2.step.item())


def test_train_tail_is_objective_step_and_scaled_adam():
    """dpgp_train_tail: *objective = scal[0] - scal[1] - gp, ++step, and Adam on g_scale[i] * grads[i] * sigmoid(raws[i]) --
    against the same update done with torch ops + dpgp_adam_multi."""
    from dp_gp_lvm_b200.engine import BoundEngine, MODE_T
    eng = BoundEngine(8, 3, 2, 4, 2, MODE_T, device=DEV)
    gen = torch.Generator(device=DEV); gen.manual_seed(9)
    shapes = [(37, 3), (), (5,), (1000,)]
    mk = lambda: [torch.randn(s, dtype=torch.float64, device=DEV, generator=gen) for s in shapes]
    th_a = mk(); th_b = [t.clone() for t in th_a]
    m_a = [torch.zeros_like(t) for t in th_a]; v_a = [torch.zeros_like(t) for t in th_a]
    m_b = [torch.zeros_like(t) for t in th_a]; v_b = [torch.zeros_like(t) for t in th_a]
    step_a = torch.zeros((), dtype=torch.int64, device=DEV); step_b = torch.zeros((), dtype=torch.int64, device=DEV)
    raw = torch.randn(37, 3, dtype=torch.float64, device=DEV, generator=gen)
    scales = [-1.0, 1.0, -1.0, 0.5]; raws = [raw, None, None, None]
    obj = torch.zeros((), dtype=torch.float64, device=DEV)
    for it in range(4):
        g = mk()
        scal = torch.randn(2, dtype=torch.float64, device=DEV, generator=gen); gp = torch.randn(1, dtype=torch.float64, device=DEV, generator=gen)
        eng.train_tail(scal, gp, obj, th_a, g, m_a, v_a, scales, raws, step_a, 0.05)
        step_b += 1
        gb = [sc * (gg * torch.sigmoid(r) if r is not None else gg) for sc, gg, r in zip(scales, g, raws)]
        eng.adam_multi(th_b, gb, m_b, v_b, step_b, 0.05)
        assert abs(float(obj.item()) - float((scal[0] - scal[1] - gp[0]).item())) == 0.0
        assert int(step_a.item()) == it + 1
    for a, b in zip(th_a + m_a + v_a, th_b + m_b + v_b):
        den = max(float(b.abs().max()), 1e-300)
        assert float((a - b).abs().max()) / den < 1e-14
