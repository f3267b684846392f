"""The caller of the hot path: the optimiser step.

Every reference script trains with `tf.train.AdamOptimizer(learning_rate).minimize(loss=model.objective)` and then
calls `session.run(train_op)` in a loop (test/synthetic_data_hard_test.py:143-155, test/cmu_walking_tests.py:143-154).
The same two steps here:

    train_op = AdamOptimizer(learning_rate=0.01).minimize(loss=model)      # model: a Trainable
    for c in range(train_iter):
        train_op.run()

One `run()` = objective + all gradients (the CUDA hot path through include/dpgp.h) + one Adam update of every
trainable variable by `dpgp_adam_multi` (one launch; TensorFlow-1's formulation, so trajectories are comparable with the reference).
With `use_cuda_graph=True` the whole iteration -- the fused small-variable kernels (dpgp_small_fwd / _bwd), the dpgp_*
launches, the NCCL all-reduces and the Adam updates -- is captured once and replayed, which removes the host
launch overhead that dominates the small configurations.
"""
import gc

import torch


class TrainOp:
    def __init__(self, model, learning_rate, beta1, beta2, epsilon, var_list=None, use_cuda_graph=False, objective_fn=None,
                 check_every=None):
        self.model = model
        self.params = list(model.parameters()) if var_list is None else list(var_list)
        self.engine = model.engine
        self.lr, self.beta1, self.beta2, self.eps = float(learning_rate), float(beta1), float(beta2), float(epsilon)
        # the model's own variables and its own objective: the whole iteration can go through the C ABI without autograd
        # (models/dp_gp_lvm.py: fused_adam_iteration); anything else takes the autograd path below
        self._fast = objective_fn is None and var_list is None and hasattr(model, "fused_adam_iteration")
        self.objective_fn = objective_fn if objective_fn is not None else (lambda: model.objective)
        dev = self.params[0].device
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.step = torch.zeros((), dtype=torch.int64, device=dev)
        self.last_objective = torch.zeros((), dtype=torch.float64, device=dev)
        self.use_cuda_graph = bool(use_cuda_graph)
        self._graph = None
        self.iterations = 0
        # tf.cholesky aborts the session at the iteration where K_uu + 1e-8 I or beta H + I stops being positive definite.
        # Here the factorisation flags the failure on the device and turns the objective and every gradient of that
        # evaluation into NaN; reading the flag needs a stream synchronisation, so it is polled every `check_every`
        # iterations: every iteration when launching eagerly (default 1), every 50 replays of the CUDA graph (default 50).
        self.check_every = (50 if self.use_cuda_graph else 1) if check_every is None else int(check_every)

    def _iteration(self):
        if self._fast and self.model.fused_adam_iteration(self.params, self.m, self.v, self.step, self.last_objective, self.lr,
                                                          self.beta1, self.beta2, self.eps):
            return
        obj = self.objective_fn()
        grads = torch.autograd.grad(obj, self.params, allow_unused=True)
        self.step += 1
        sel = [(p.data, g.contiguous(), m, v) for p, g, m, v in zip(self.params, grads, self.m, self.v) if g is not None]
        if hasattr(self.engine, "adam_multi"):
            self.engine.adam_multi([a[0] for a in sel], [a[1] for a in sel], [a[2] for a in sel], [a[3] for a in sel], self.step,
                                   self.lr, self.beta1, self.beta2, self.eps)
        else:
            for p, g, m, v in sel:
                self.engine.adam(p, g, m, v, self.step, self.lr, self.beta1, self.beta2, self.eps)
        self.last_objective.copy_(obj.detach())

    def _capture(self):
        # Warm-up iterations on a side stream (allocator, lazy initialisations) as torch.cuda.graph requires; the
        # optimiser state is restored afterwards so that run() k times == k iterations, with or without the graph.
        saved = [t.detach().clone() for t in self.params + self.m + self.v] + [self.step.clone()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._iteration()
        torch.cuda.current_stream().wait_stream(s)
        with torch.no_grad():
            for t, c in zip(self.params + self.m + self.v + [self.step], saved):
                t.copy_(c)
        # Python's cycle collector must not run inside the capture: finalisers of dead CUDA objects (another model's engine, an
        # old CUDAGraph) free device memory or synchronise streams, which invalidates a capture in progress.  Collect now, keep
        # the collector off until the capture has ended.
        gc.collect()
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._iteration()
        except Exception:
            self._graph = None
            raise
        finally:
            if gc_was_enabled:
                gc.enable()

    def run(self):
        """One optimisation iteration (the analogue of session.run(train_op))."""
        if not self.use_cuda_graph:
            self._iteration()
        else:
            if self._graph is None:
                self._capture()
            self._graph.replay()
        self.iterations += 1
        if self.check_every > 0 and self.iterations % self.check_every == 0 and hasattr(self.engine, "check"):
            self.engine.check()              # raises NotPositiveDefiniteError with the pivot location

    @property
    def objective(self):
        """Objective evaluated at the start of the most recent iteration (device scalar; no extra evaluation)."""
        return self.last_objective


class AdamOptimizer:
    """Same constructor arguments and defaults as tf.train.AdamOptimizer."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-08, use_cuda_graph=False, check_every=None):
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon
        self.use_cuda_graph = use_cuda_graph
        self.check_every = check_every

    def minimize(self, loss, var_list=None, objective_fn=None):
        """`loss` is the Trainable model (its `.objective` is re-evaluated every iteration, as a TF-1 graph node is on
        every session.run); `var_list` defaults to all of its trainable variables (tf's TRAINABLE_VARIABLES)."""
        return TrainOp(loss, self.learning_rate, self.beta1, self.beta2, self.epsilon, var_list=var_list,
                       use_cuda_graph=self.use_cuda_graph, objective_fn=objective_fn, check_every=self.check_every)
