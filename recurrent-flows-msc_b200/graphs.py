"""CUDA-graph capture of the hot path.

One GlowStep is 5 small launches and ListGlow has 50 of them, so the Python / ctypes cost of enqueueing (~20 us per
launch) rivals the GPU time of a whole call, and dominates the autoregressive sampling path where every predicted
frame runs ListGlow.sample on only B sequences.  ``Graphed`` captures one call of a function of
tensors for fixed shapes into a CUDA graph -- kernel launches go through the C ABI on torch's capture stream; TMA
descriptors are baked in as kernel parameters, which is valid because staging buffers come from the persistent
workspace pool and outputs are graph-private -- and replays it with new inputs copied into static buffers.
Parameters must not change between capture and replay (re-capture after an optimizer step / load_state_dict).
"""
import torch


def _map(x, f):
    if torch.is_tensor(x):
        return f(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_map(v, f) for v in x)
    return x


def _copy_into(dst, src):
    if torch.is_tensor(dst):
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            _copy_into(d, s)


class Graphed:
    """Graphed(fn, *example_inputs)(*inputs) -> fn's outputs, as views of static buffers (clone to keep them)."""

    def __init__(self, fn, *example_inputs, warmup=2):
        self.static_in = _map(list(example_inputs), lambda t: t.detach().clone())
        with torch.no_grad():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):   # builds weight caches, workspaces, ActNorm flags outside the capture
                    fn(*self.static_in)
            torch.cuda.current_stream().wait_stream(s)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = fn(*self.static_in)
        from . import ops
        ops.register_graph(self)

    def __call__(self, *inputs):
        _copy_into(self.static_in, list(inputs))
        self.graph.replay()
        return self.out


class GraphedSample(Graphed):
    """Graph of ``flow.sample(None, condition, base_condition, temperature=...)`` (the RFN.predict inner step)."""

    def __init__(self, flow, condition, base_condition, temperature=0.8, warmup=2):
        n = condition[-1].shape[0]
        if base_condition is None:
            super().__init__(lambda c: flow.sample(None, c, None, num_samples=n, temperature=temperature),
                             list(condition), warmup=warmup)
        else:
            super().__init__(lambda c, b: flow.sample(None, c, b, num_samples=n, temperature=temperature),
                             list(condition), base_condition, warmup=warmup)
        self._has_base = base_condition is not None

    def __call__(self, condition, base_condition=None):
        return super().__call__(list(condition), base_condition) if self._has_base else super().__call__(list(condition))


class GraphedLogProb(Graphed):
    """Graph of ``flow.log_prob(x, condition, base_condition)`` -> (z, nll); the dequantisation draw is graph-safe."""

    def __init__(self, flow, x, condition, base_condition, warmup=2):
        super().__init__(lambda xx, c, b: flow.log_prob(xx, c, b), x, list(condition), base_condition, warmup=warmup)

    def __call__(self, x, condition, base_condition):
        return super().__call__(x, list(condition), base_condition)


class GraphedTrainStep:
    """One whole training step -- zero_grad, forward with tape, hand-written backward, gradient gather, (allreduce),
    fused Adam, and the weight repacking the next forward needs -- captured into CUDA graphs and replayed.

    ``loss_fn()`` must compute the loss from STATIC input tensors (copy new batches into them before calling) and
    return it; ``optimizer`` is a ``FlatAdam``.  With more than one replica the NCCL allreduce runs eagerly between two
    graphs (forward+backward+gather | Adam), so nothing depends on collective capture support.
    """

    def __init__(self, loss_fn, optimizer, warmup=3, static_inputs=None):
        from .Flow.glow_modules import invalidate_caches
        self._invalidate = invalidate_caches
        self.opt = optimizer
        self.static_inputs = static_inputs   # optional (nested) list of the tensors loss_fn reads: step(*batch) copies into them

        def fwd_bwd():
            optimizer.zero_grad()      # one memset of the flat gradient buffer; .grad tensors stay views of it
            loss = loss_fn()
            loss.backward()
            with torch.no_grad():
                optimizer.gather_grads()
            return loss.detach()

        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):   # ActNorm init, function attributes, allocator warm-up; real optimizer steps
                fwd_bwd()
                optimizer.allreduce_grads()
                with torch.no_grad():
                    optimizer.apply()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # ONE graph for the whole step, the NCCL all-reduce of the flat gradient included (captured like any other stream
        # work: no host round trip between backward, collective and Adam).  RFK_GRAPH_ALLREDUCE=0, or a failed capture
        # of the collective, falls back to two graphs around an eager all-reduce.
        import os
        self.g_all = self.g_fb = self.g_opt = None
        if optimizer.world == 1 or os.environ.get("RFK_GRAPH_ALLREDUCE", "1") != "0":
            try:
                g = torch.cuda.CUDAGraph()
                # thread_local: the NCCL watchdog thread may touch the CUDA API (event queries) while this thread captures
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self.loss = fwd_bwd()
                    optimizer.allreduce_grads()
                    with torch.no_grad():
                        optimizer.apply()
                self.g_all = g
            except Exception as e:  # noqa: BLE001
                if optimizer.world == 1:
                    raise
                import warnings
                warnings.warn(f"recurrent-flows-msc_b200: capturing the NCCL all-reduce failed ({e!r}); using two graphs around an eager all-reduce")
                torch.cuda.synchronize()
        if self.g_all is None:
            self.g_fb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_fb, capture_error_mode="thread_local"):
                self.loss = fwd_bwd()
            self.g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g_opt, capture_error_mode="thread_local"), torch.no_grad():
                optimizer.apply()
        from . import ops
        ops.register_graph(self)
        self.mode = "one CUDA graph (forward + backward + all-reduce + Adam)" if self.g_all is not None else \
            "two CUDA graphs (forward + backward | Adam) around an eager all-reduce"

    def __call__(self, *batch):
        if batch:
            if self.static_inputs is None:
                raise ValueError("GraphedTrainStep: pass static_inputs= at construction to feed new batches")
            _copy_into(self.static_inputs, list(batch))
        if self.g_all is not None:
            self.g_all.replay()
        else:
            self.g_fb.replay()
            self.opt.allreduce_grads()
            self.g_opt.replay()
        self._invalidate()   # eager calls after a replay must not trust caches filled before the last update ...
        from .derived import REFRESHER
        REFRESHER.restamp()  # ... except the registered ones: the optimizer graph has just rewritten them in place
        return self.loss
