#!/usr/bin/env python
"""bench.py -- DP-GP-LVM ELBO+gradient evaluations per second at the BASELINE.json headline shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rows N_TOTAL]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one evaluation of the T-mode objective AND its gradients w.r.t. every trainable variable
(no optimiser step) on synthetic data of the headline shape N = 1,000,000, D = 64, Q = 10, M = 128, T = 10
(`configs[4]`), float64.  With N ranks the N rows are sharded (strong scaling: the job is fixed), the
packed statistics are all-reduced over NCCL, and the time is the max over ranks of the device time.

JSON keys beyond the base contract:
  value     evals/s with every input resident in HBM (CUDA events around K steps)
  e2e       the same through the public model API with HOST buffers: every step copies this rank's rows of Y
            and all trainable variables host->device from pinned memory and copies the objective and ALL
            gradients back (the reference feeds y_train once as a graph constant, dp_gp_lvm.py:143,657; it is
            re-sent here every step so that no input of the timed region is device-resident).  The copies are
            PIPELINED with the compute of the neighbouring steps on two side streams (inputs of step k+1 go to
            staging buffers while step k computes; the gradients of step k leave while step k+1 computes); every
            byte is still copied every step and the timed region ends only when the last result is in host memory.
  roofline  the dominant kernel of the step: algorithmic flops (71 flops per (cluster, n, m<=m') unit
            for the psi2 forward kernel, 152 for the fused psi2 backward kernel at Q = 10) / measured launch time,
            against the FP64 pipe peak measured on this pool (profiles/r01_fp64_peaks.json; MEASURED_PEAKS.json
            has no FP64 figure).  `kernels` lists both psi2 kernels.
  cpu_baseline  the CPU oracle port (oracle/streaming.py, torch float64) on a bounded row sample of the same
            workload, with the linearity residual between two sample sizes.
  configs   (1 GPU only) BASELINE.json configs[0..3] at their exact shapes: GPU evals/s in T- and D-mode, Adam
            iterations/s (one CUDA graph per iteration), and the literal CPU oracle (oracle/literal.py, the
            reference's op sequence with its materialised tensors) timed as it is for configs[0..2].
`--impl reference` times the CPU port alone on the same workload and the same kind of sample (TensorFlow 1.15
cannot be installed here, so the oracle port is the reference arm).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = dict(n=1_000_000, d=64, q=10, m=128, t=10)
FP64_PEAK_TFLOPS = 37.1           # measured DMMA/DFMA pipe capacity, profiles/r01_fp64_peaks.json
METRIC = "DP-GP-LVM ELBO+grad evals/s at N=1M"
# BASELINE.json configs[0..3] at their exact shapes: (name, source, n, d, q, m, t, mask)
CONFIGS = [
    ("configs[0]", "synthetic_data_test.py small synthetic", 100, 10, 2, 25, 8, 1),
    ("configs[1]", "synthetic_data_hard_test.py", 100, 60, 10, 50, 20, 1),
    ("configs[2]", "CMU walking shape", 300, 60, 10, 50, 10, 3),
    ("configs[3]", "Frey faces shape", 1965, 560, 10, 100, 20, 1),
]


def workload_string(shape):
    """Identical for both arms (the driver compares it)."""
    return "configs[4]: N=%d D=%d Q=%d M=%d T=%d, T-mode ELBO+grad" % (shape["n"], shape["d"], shape["q"], shape["m"], shape["t"])


def fp64_peak():
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")) as f:
            return float(json.load(f)["dmma_m8n8k4_tflops"]), "measured (profiles/r01_fp64_peaks.json, DMMA m8n8k4 = FP64 pipe capacity)"
    except Exception:
        return FP64_PEAK_TFLOPS, "fallback constant"


def synthetic(n_rows, row0, shape, seed=0):
    """Rows [row0, row0+n_rows) of the synthetic problem; identical regardless of how rows are sharded."""
    import numpy as np
    d, q, m, t = shape["d"], shape["q"], shape["m"], shape["t"]
    mask = shape.get("mask", 1)
    rng = np.random.default_rng(seed)
    base = float(np.log(np.expm1(1.0)))
    small = dict(
        x_u=rng.standard_normal((m, q)), phi_logits=rng.standard_normal((d // mask, t)),
        gamma1_raw=rng.standard_normal(t - 1), gamma2_raw=rng.standard_normal(t - 1),
        w1_raw=np.array(base), w2_raw=np.array(base),
        gamma_atoms_raw=base + 0.3 * rng.standard_normal((t, q)), alpha_atoms_raw=base + 0.3 * rng.standard_normal((t, 1)),
        beta_atoms_raw=base + 0.3 * rng.standard_normal((t, 1)))
    # per-row streams keyed by the global row block so shards agree with the single-GPU problem
    blk = 65536
    ys, ms, ss = [], [], []
    b0, b1 = row0 // blk, (row0 + n_rows - 1) // blk
    for b in range(b0, b1 + 1):
        r = np.random.default_rng([seed, 1000 + b])
        yb = r.standard_normal((blk, d)); mb = r.standard_normal((blk, q)); sb = base + 0.1 * r.standard_normal((blk, q))
        lo = max(row0, b * blk) - b * blk; hi = min(row0 + n_rows, (b + 1) * blk) - b * blk
        ys.append(yb[lo:hi]); ms.append(mb[lo:hi]); ss.append(sb[lo:hi])
    return np.concatenate(ys), dict(x_mean=np.concatenate(ms), x_var_raw=np.concatenate(ss), **small)


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index; self.rows = []; self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) > 3 + i and r[3 + i].lower() == "active" for r in self.rows):
                reasons.append(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ CPU legs (oracle/)
def cpu_port_eval(shape, n_sample, threads, reps=1):
    """One ELBO+grad evaluation of the CPU oracle port on n_sample rows; returns seconds per evaluation."""
    import torch
    from oracle import streaming as S
    torch.set_num_threads(threads)
    y, params = synthetic(n_sample, 0, shape)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        S.value_and_grad(y, params, "t", chunk=128)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_sample_rows(shape, threads, budget_s, evaluations, cap=4096):
    """Rows of the bounded CPU sample: the largest multiple of 256 (<= cap, the size SURVEY.md 8d asks for) such that
    `evaluations` evaluations fit in budget_s, from a 256-row calibration run.  Returns (rows, seconds of the calibration)."""
    t256 = cpu_port_eval(shape, 256, threads)
    rows = int(256 * budget_s / max(evaluations * t256, 1e-9)) // 256 * 256
    return max(256, min(cap, rows, shape["n"])), t256


def cpu_baseline_block(shape, threads, rows):
    """The CPU port on `rows` rows and on rows / 2 (linearity residual), scaled linearly to the full N."""
    n = shape["n"]
    dt = cpu_port_eval(shape, rows, threads)
    half = max(128, rows // 2)
    dt_half = cpu_port_eval(shape, half, threads)
    full = dt * n / rows
    return {"value": 1.0 / full, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": "%d of %d rows, one evaluation of oracle/streaming.py (torch float64 CPU), scaled linearly in N" % (rows, n),
            "sample_rows": rows, "sample_seconds": dt,
            "linearity": {"rows": [half, rows], "seconds": [dt_half, dt],
                          "residual": abs(dt - dt_half * rows / half) / dt,
                          "note": "every N-dependent term is a sum over rows; residual = |t(rows) - t(rows/2) * 2| / t(rows)"}}


def literal_cpu_eval(name, n, d, q, m, t, mask, mode, threads, seed):
    """configs[0..2] through the LITERAL CPU oracle (the reference's op sequence, [B,N,M,M,Q] tensors and all), timed as is."""
    import torch
    from oracle import literal as L
    torch.set_num_threads(threads)
    b = t if mode == "t" else d
    need_gb = 8.0 * b * n * m * m * q * 8 / 1e9              # ~8 live [B,N,M,M,Q] tensors under autograd
    try:
        import psutil
        avail = psutil.virtual_memory().available / 1e9
    except Exception:
        avail = 16.0
    if need_gb > 0.6 * avail:
        return {"skipped": "needs ~%.0f GB for the materialised [B,N,M,M,Q] tensors (%.0f GB available)" % (need_gb, avail)}
    y, params = synthetic(n, 0, dict(d=d, q=q, m=m, t=t, mask=mask), seed=seed)
    fn = L.objective_t if mode == "t" else L.objective_d
    t0 = time.perf_counter()
    obj, _ = L.value_and_grad(fn, y, params, (1.0, 1.0), mask)
    dt = time.perf_counter() - t0
    return {"evals_per_s": 1.0 / dt, "seconds": dt, "objective": obj, "cores": threads}


# ------------------------------------------------------------------------------------------------ GPU side legs
def train_bench(dev, iters=60):
    """The caller of the hot path on configs[1] (synthetic_data_hard_test.py: N=100, D=60, Q=10, M=50, T=20):
    Adam iterations per second, eager and as one replayed CUDA graph per iteration.  Not the headline metric."""
    import numpy as np
    import torch
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm
    from dp_gp_lvm_b200.train import AdamOptimizer
    rng = np.random.default_rng(10)
    y = rng.standard_normal((100, 60))
    res = {"config": "configs[1]: N=100 D=60 Q=10 M=50 T=20, D-mode, Adam lr=0.01", "iterations": iters}
    for key, graph in (("iters_per_s_eager", False), ("iters_per_s_cuda_graph", True)):
        np.random.seed(10)
        model = dp_gp_lvm(y_train=y, num_latent_dims=10, num_inducing_points=50, truncation_level=20, device=dev)
        op = AdamOptimizer(learning_rate=0.01, use_cuda_graph=graph).minimize(loss=model)
        for _ in range(5):
            op.run()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            op.run()
        torch.cuda.synchronize()
        res[key] = iters / (time.perf_counter() - t0)
        res["objective_after_%s" % ("graph" if graph else "eager")] = float(op.objective.item())
        model.engine.check()
    return res


def configs_block(dev, threads, with_cpu=True, evals=30, iters=60):
    """BASELINE.json configs[0..3] at their exact shapes on one GPU (and the literal CPU oracle for configs[0..2])."""
    import numpy as np
    import torch
    from dp_gp_lvm_b200.models.dp_gp_lvm import dp_gp_lvm, dp_gp_lvm_t
    from dp_gp_lvm_b200.train import AdamOptimizer
    out = []
    for idx, (name, src, n, d, q, m, t, mask) in enumerate(CONFIGS):
        shape = dict(n=n, d=d, q=q, m=m, t=t, mask=mask)
        y, params = synthetic(n, 0, shape, seed=100 + idx)
        entry = {"config": "%s (%s): N=%d D=%d Q=%d M=%d T=%d mask=%d" % (name, src, n, d, q, m, t, mask)}
        for mode in ("t", "d"):
            print("[bench] %s %s-mode" % (name, mode), file=sys.stderr, flush=True)
            kw = dict(y_train=y, num_latent_dims=q, num_inducing_points=m, truncation_level=t, mask_size=mask, device=dev)

            def build():
                np.random.seed(0)
                mdl = dp_gp_lvm_t(seed=0, **kw) if mode == "t" else dp_gp_lvm(**kw)
                mdl.load_variables(params)
                return mdl
            # evaluations per second, eager launches (objective + all gradients, no optimiser step)
            model = build()
            leaves = model.parameters()

            def step():
                obj = model.objective
                torch.autograd.grad(obj, leaves, allow_unused=True)
                return obj
            for _ in range(5):
                step()
            model.engine.check()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(evals if n * (t if mode == "t" else d) < 500000 else 10):
                obj = step()
            e1.record()
            torch.cuda.synchronize()
            done = evals if n * (t if mode == "t" else d) < 500000 else 10
            res = {"gpu_evals_per_s": done / (e0.elapsed_time(e1) * 1e-3), "objective": float(obj.item())}
            del model, leaves, obj
            # Adam iterations as one replayed CUDA graph each (what the training loop does), on a fresh model: autograd nodes
            # created on the legacy default stream by the eager evaluations above cannot be re-used under capture
            model = build()
            op = AdamOptimizer(learning_rate=0.01, use_cuda_graph=True).minimize(loss=model)
            for _ in range(3):
                op.run()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                op.run()
            torch.cuda.synchronize()
            res["adam_iters_per_s_cuda_graph"] = iters / (time.perf_counter() - t0)
            model.engine.check()
            if with_cpu and idx < 3:
                cpu = literal_cpu_eval(name, n, d, q, m, t, mask, mode, threads, 100 + idx)
                res["cpu_literal_oracle"] = cpu
                if "evals_per_s" in cpu:
                    res["objective_rel_diff_vs_cpu"] = abs(res["objective"] - cpu["objective"]) / abs(cpu["objective"])
            entry["t_mode" if mode == "t" else "d_mode"] = res
            del model, op
            torch.cuda.empty_cache()
        out.append(entry)
    return out


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, shape):
    """Reference arm: the CPU float64 port of the reference graph on the host cores (rank 0 only).  Every step is one
    evaluation on a bounded row sample of the same workload, sized so that the K timed steps take ~3 minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    if args.cpu_rows > 0:
        n_sample = args.cpu_rows
    else:
        n_sample, _ = cpu_sample_rows(shape, threads, 180.0, max(args.steps, 1))
    for _ in range(args.warmup):
        cpu_port_eval(shape, min(n_sample, 64), threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_eval(shape, n_sample, threads)
    dt = (time.perf_counter() - t0) / args.steps
    full = dt * shape["n"] / n_sample          # linear in N (every N-dependent term is a sum over rows)
    val = 1.0 / full
    sample = "%d of %d rows per step (chunked streaming oracle, torch float64 CPU), time scaled linearly in N" % (n_sample, shape["n"])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(shape)},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port", "sample": sample, "sample_rows": n_sample,
                         "sample_seconds": dt},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rows", type=int, default=SHAPE["n"], help="total rows N (default: the headline 1,000,000)")
    ap.add_argument("--cpu-rows", type=int, default=0, help="rows of the bounded CPU sample (0: as many as fit the time budget, <= 4096)")
    ap.add_argument("--exp-variant", type=int, default=0)
    ap.add_argument("--bwd-variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the small training-loop measurement")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0..3] block")
    args = ap.parse_args()
    shape = dict(SHAPE, n=args.rows)
    if args.impl == "reference":
        return run_reference(args, shape)

    import numpy as np
    import torch
    import torch.distributed as dist
    from dp_gp_lvm_b200.models.dp_gp_lvm import PARAM_ORDER, dp_gp_lvm_t

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        # keep stdout to the one JSON line: this image exports NCCL_DEBUG=VERSION and NCCL prints its banner on stdout
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # the banner still reached stdout on the 2 / 4 / 8-GPU runs of round 2
            os.environ["NCCL_DEBUG"] = "WARN"
        # belt and braces: whatever the libraries print while the communicator comes up goes to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            group = dist.group.WORLD
            dist.all_reduce(torch.zeros(1, device=dev))         # creates the communicator now
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    assert world == args.gpus, "launch with torchrun --nproc-per-node == --gpus"

    n = shape["n"]
    lo, hi = n * rank // world, n * (rank + 1) // world
    y, params = synthetic(hi - lo, lo, shape)
    # model through the public API (the factory's own PCA initialisation is overwritten by the synthetic point)
    model = dp_gp_lvm_t(y_train=y, num_latent_dims=shape["q"], num_inducing_points=shape["m"], truncation_level=shape["t"],
                        seed=0, device=dev, process_group=group, exp_variant=args.exp_variant, bwd_variant=args.bwd_variant)
    model.load_variables(params)
    leaves = model.parameters()
    eng = model.engine
    eng.set_timing(True)

    def step():
        for p in leaves:
            p.grad = None
        obj = model.objective
        obj.backward()
        return obj

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    eng.check()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    phases = eng.timings()                # last step's per-phase device times (ms)
    # torch-side kernels (softplus/softmax/DP objective/all-reduce) are not counted in gpu_launches
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    obj_val = float(step().item())
    clocks = sampler.stop() if rank == 0 else None
    eng.set_timing(False)

    # ---- end to end: host-resident variables in, objective + all gradients out, every step, copies pipelined with compute
    host_in = [p.detach().cpu().pin_memory() for p in leaves]
    host_out = [torch.empty_like(h).pin_memory() for h in host_in]
    host_obj = torch.empty((), dtype=torch.float64).pin_memory()
    host_y = torch.from_numpy(y).pin_memory()
    y_dev = model.y_train_device
    d2h = sum(h.numel() * 8 for h in host_in) + 8
    h2d = sum(h.numel() * 8 for h in host_in) + host_y.numel() * 8
    main_s = torch.cuda.current_stream()
    in_s, out_s = torch.cuda.Stream(), torch.cuda.Stream()
    stage_p = [torch.empty_like(p.detach()) for p in leaves]       # device staging: inputs of the NEXT step land here
    stage_y = torch.empty_like(y_dev)
    ev_staged, ev_consumed = torch.cuda.Event(), torch.cuda.Event()

    def stage_inputs():
        """Host -> device staging on the input stream (after the previous contents were consumed)."""
        in_s.wait_event(ev_consumed)
        with torch.cuda.stream(in_s):
            stage_y.copy_(host_y, non_blocking=True)
            for s_, h in zip(stage_p, host_in):
                s_.copy_(h, non_blocking=True)
            ev_staged.record(in_s)

    def e2e_step():
        main_s.wait_event(ev_staged)
        with torch.no_grad():
            y_dev.copy_(stage_y)                                    # device -> device, ~0.3 ms for 672 MB
            for p, s_ in zip(leaves, stage_p):
                p.copy_(s_)
        ev_consumed.record(main_s)
        stage_inputs()                                              # inputs of the next step travel while this one computes
        obj = step()
        ev_done = torch.cuda.Event(); ev_done.record(main_s)
        out_s.wait_event(ev_done)
        with torch.cuda.stream(out_s):
            host_obj.copy_(obj.detach(), non_blocking=True)
            obj.record_stream(out_s)
            for p, h in zip(leaves, host_out):
                h.copy_(p.grad, non_blocking=True)
                p.grad.record_stream(out_s)

    ev_consumed.record(main_s)
    stage_inputs()
    e2e_step()                                                      # warm-up of the pipeline
    out_s.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    out_s.synchronize(); main_s.synchronize()                       # the last step's results are in host memory
    wall_ms = (time.perf_counter() - t0) * 1e3
    in_s.synchronize()
    barrier()
    t = torch.tensor([wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    e2e_obj = float(host_obj.item())

    if rank == 0:
        peak, peak_src = fp64_peak()
        b, m, q = shape["t"], shape["m"], shape["q"]
        units = b * (hi - lo) * (m * (m + 1) // 2)            # per rank, per launch
        f_fwd = units * (4 * q + 3 + 28)
        f_bwd = units * (12 * q + 4 + 28)

        def tf(flops, ms):
            return flops / (ms * 1e-3) / 1e12 if ms and ms > 0 else None
        fwd_ms = phases.get("psi2_fwd")
        a_fwd = tf(f_fwd, fwd_ms)
        bwd_ms = (phases.get("psi2_bwd_fused", 0) or 0) + (phases.get("psi2_bwd_n", 0) or 0) + (phases.get("psi2_bwd_pair", 0) or 0)
        bwd_name = "psi2_bwd_fused_kernel" if phases.get("psi2_bwd_fused") else "psi2_bwd_n_kernel+psi2_bwd_pair_kernel"
        a_bwd = tf(f_bwd, bwd_ms)
        # DRAM traffic per launch from the committed `ncu --set full` capture of this same command (profiles/)
        traffic = {}
        for tj_name in ("r02_traffic.json", "r01_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", tj_name)) as f:
                    tj = json.load(f)
                if tj.get("rows_per_launch") == hi - lo:
                    traffic = tj.get("dram_bytes_per_launch", {})
                    break
            except Exception:
                pass
        kern = {"psi2_fwd_kernel": {"achieved": a_fwd, "frac": (a_fwd / peak) if a_fwd else None, "launch_ms": fwd_ms,
                                    "algorithmic_flops_per_launch": f_fwd, "share_of_step": fwd_ms / ms_step if fwd_ms else None,
                                    "traffic": traffic.get("psi2_fwd_kernel")},
                bwd_name: {"achieved": a_bwd, "frac": (a_bwd / peak) if a_bwd else None, "launch_ms": bwd_ms,
                           "algorithmic_flops_per_launch": f_bwd, "share_of_step": bwd_ms / ms_step if bwd_ms else None,
                           "traffic": traffic.get(bwd_name)}}
        dom = max(kern, key=lambda k: kern[k]["launch_ms"] or 0.0)          # the dominant kernel of the step
        roofline = {"bound": "tensor", "pipe": "FP64 (DFMA and DMMA share one pipe on B200: profiles/r01_fp64_peaks.md)",
                    "kernel": dom, "achieved": kern[dom]["achieved"], "peak": peak, "unit": "TFLOP/s",
                    "frac": kern[dom]["frac"], "traffic": kern[dom]["traffic"], "peak_source": peak_src,
                    "launch_ms": kern[dom]["launch_ms"], "algorithmic_flops_per_launch": kern[dom]["algorithmic_flops_per_launch"],
                    "kernels": kern, "phases_ms": phases}
        out = {
            "metric": METRIC, "value": 1e3 / ms_step, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(shape),
                       "l2": "inputs (%.0f MB/rank) and workspace (%.1f GB/rank) exceed the 126 MB L2" % ((hi - lo) * (2 * q + shape["d"]) * 8 / 1e6, eng.workspace_bytes / 1e9),
                       "parallelism": "rows sharded over %d GPU(s), all-reduce of %d doubles" % (world, eng.stats_len), "objective": obj_val},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": 1e3 / e2e_ms, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                    "objective_read_back": e2e_obj,
                    "copies": "pinned host -> device staging and device -> pinned host on two side streams, pipelined with the neighbouring steps"},
            "roofline": roofline}
        threads = os.cpu_count() or 1
        if not args.no_train and world == 1:
            out["train"] = train_bench(dev)
        if not args.no_configs and world == 1:
            out["configs"] = configs_block(dev, threads, with_cpu=not args.no_cpu_baseline)
        if not args.no_cpu_baseline and world == 1:
            rows = args.cpu_rows if args.cpu_rows > 0 else cpu_sample_rows(shape, threads, 45.0, 1)[0]
            out["cpu_baseline"] = cpu_baseline_block(shape, threads, rows)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
