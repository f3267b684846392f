"""Two ranks on two real GPUs over NCCL (skipped on a one-GPU box): the row-sharded model equals the single-GPU model.

Covers what the gloo test (tests/test_distributed_cpu.py) cannot: `_distributed_pca` and the CUDA engine under a process
group -- all-reduced statistics, replicated M x M chain, sharded q(X) gradients, all-reduced (Z, gamma, alpha) partials."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from dp_gp_lvm_b200.models.dp_gp_lvm import PARAM_ORDER, dp_gp_lvm_t
    from dp_gp_lvm_b200.utils.expressions import principal_component_analysis as pca
    from oracle.literal import random_params
    rng = np.random.default_rng(5)
    n, d, q, m, t = 301, 12, 4, 20, 5                        # odd N: unequal shards
    y = rng.standard_normal((n, d))
    params = random_params(rng, n, d, q, m, t)
    lo, hi = n * rank // world, n * (rank + 1) // world
    model = dp_gp_lvm_t(y_train=y[lo:hi], num_latent_dims=q, num_inducing_points=m, truncation_level=t, seed=0, device=dev,
                        process_group=dist.group.WORLD)
    # (1) distributed PCA initialisation == PCA of the whole matrix, up to column signs
    x0 = model.variables["x_mean"].detach().cpu().numpy()
    full = pca(y, q)[lo:hi]
    s = np.sign((x0 * full).sum(axis=0))
    pca_err = float(np.abs(x0 * s - full).max() / np.abs(full).max())
    # (2) replicated variables start identical on every rank
    xu = model.variables["x_u"].detach().clone()
    ref = xu.clone(); dist.broadcast(ref, src=0)
    same_xu = bool(torch.equal(xu, ref))
    # (3) objective and gradients at a common point
    local = dict(params); local["x_mean"] = params["x_mean"][lo:hi]; local["x_var_raw"] = params["x_var_raw"][lo:hi]
    model.load_variables(local)
    obj, grads = model.value_and_grad()
    res = {"obj": obj, "pca_err": pca_err, "same_xu": same_xu, "n_total": model.num_samples_total,
           "grads": {k: grads[k] for k in PARAM_ORDER}, "lo": lo, "hi": hi}
    if rank == 0:
        single = dp_gp_lvm_t(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, seed=0, device=dev)
        single.load_variables(params)
        res["single"] = single.value_and_grad()
    out.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_model_equals_single_gpu_model():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    r0, r1 = got[0], got[1]
    obj1, g1 = r0["single"]
    assert r0["n_total"] == r1["n_total"] == 301
    assert r0["pca_err"] < 1e-9 and r1["pca_err"] < 1e-9
    assert r0["same_xu"] and r1["same_xu"]
    assert r0["obj"] == r1["obj"], "the replicated chain must be bit-identical across ranks"
    assert abs(r0["obj"] - obj1) <= 1e-12 * abs(obj1)
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    for k, g in g1.items():
        if k in ("x_mean", "x_var_raw"):
            both = np.concatenate([r0["grads"][k], r1["grads"][k]], axis=0)
            assert rel(both, g) < 1e-11, k
        else:
            assert np.array_equal(r0["grads"][k], r1["grads"][k]), k
            assert rel(r0["grads"][k], g) < 1e-10, k
