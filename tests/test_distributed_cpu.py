"""world_size-2 gloo test (CPU) of the N-sharded path: each rank holds half of the rows of Y and of q(X);
the packed statistics and the small gradient partials are all-reduced exactly as on NCCL.  The CUDA engine is
replaced by the oracle-backed stand-in (tests/fake_engine.py) because there is no GPU here; what is under test
is the model's orchestration: which buffers are reduced, that sharded gradients stay local and replicated
gradients come out identical on every rank, and that the result equals the single-process objective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_params, load_golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, mode, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    import dp_gp_lvm_b200.models.dp_gp_lvm as M
    from fake_engine import OracleEngine
    M.ENGINE_FACTORY = OracleEngine
    import dp_gp_lvm_b200.utils.special as SP
    from fake_engine import scipy_polygamma
    SP.POLYGAMMA_HOOK = scipy_polygamma
    z = load_golden("%s_%s" % (mode, case))
    p = golden_params(z)
    n = z["y"].shape[0]
    lo, hi = n * rank // world, n * (rank + 1) // world
    local = dict(p); local["x_mean"] = p["x_mean"][lo:hi]; local["x_var_raw"] = p["x_var_raw"][lo:hi]
    t = p["gamma_atoms_raw"].shape[0]
    np.random.seed(0)
    kw = dict(y_train=z["y"][lo:hi], num_latent_dims=p["x_mean"].shape[1], num_inducing_points=p["x_u"].shape[0],
              truncation_level=t, alpha_prior_params=z["alpha_prior"], mask_size=int(z["mask_size"]), device="cpu",
              process_group=dist.group.WORLD)
    model = M.dp_gp_lvm_t(seed=0, **kw) if mode == "t" else M.dp_gp_lvm(**kw)
    assert model.num_samples_total == n
    model.load_variables(local)
    try:
        obj, grads = model.value_and_grad()
        q.put((rank, lo, hi, obj, grads))
    except Exception as e:      # surface the failure instead of letting the parent time out
        q.put((rank, lo, hi, repr(e), None))
        raise
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode,case", [("t", "unit"), ("d", "mask3"), ("t", "q10")])
def test_two_rank_sharded_equals_reference(mode, case):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    assert all(r[4] is not None for r in res), [r[3] for r in res]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = load_golden("%s_%s" % (mode, case))
    ref = float(z["objective"])
    res.sort(key=lambda r: r[0])
    for rank, lo, hi, obj, grads in res:
        assert abs(obj - ref) <= 1e-10 * abs(ref)
        for k, g in grads.items():
            want = z["g_" + k]
            if k in ("x_mean", "x_var_raw"):
                want = want[lo:hi]
            if want.size:
                err = np.abs(g - want).max() / max(np.abs(z["g_" + k]).max(), 1e-300)
                assert err < 1e-9, (rank, k, err)
    # replicated gradients are bitwise identical across ranks (every rank runs the same chain on the same reduced statistics)
    for k in res[0][4]:
        if k not in ("x_mean", "x_var_raw"):
            assert np.array_equal(res[0][4][k], res[1][4][k]), k
