"""Optimizer step of the training path: Adam over every parameter in ONE kernel launch, plus the data-parallel
gradient allreduce (one NCCL call on the flat gradient buffer).

The reference trains with ``torch.optim.Adam`` (RFN/trainer.py); for ListGlow's ~750 parameter tensors that is a
multi-tensor foreach pass.  ``FlatAdam`` re-homes the parameters as views of one flat fp32 buffer, gathers the gradients
into a second one (one ``torch.cat``), optionally sum-allreduces it over the process group, and runs ``rfk_adam_step``
(include/rfk.h) on the flat buffers: 28 bytes of HBM traffic per parameter, no per-tensor launches.
"""
import torch

from . import derived, ops
from ._lib import call
from .Flow.glow_modules import invalidate_caches


class FlatAdam:
    """Adam with torch.optim.Adam's defaults and update rule (no weight decay, no amsgrad).  Build it AFTER moving the
    model to its device: the parameters become views of ``self.flat_p``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, process_group=None, world_size=1):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("recurrent-flows-msc_b200: FlatAdam needs CUDA parameters (there is no CPU path)")
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        # torch.optim-style handle for LR schedulers (RFN/trainer.py:100,200 write param_groups[0]['lr'])
        self.param_groups = [{"params": self.params, "lr": float(lr), "betas": self.betas, "eps": self.eps}]
        self.group, self.world = process_group, int(world_size)
        self.overlap = True      # attach(): all-reduce per-level slices during the backward sweep (False: one call at the end)
        self.n = sum(p.numel() for p in self.params)                       # real parameters
        offs, off = [], 0
        for p in self.params:                                               # every tensor starts on a 16-byte boundary
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.n_pad = off                                                    # flat length incl. alignment padding (stays zero)
        self.flat_p = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(self.n_pad, device=dev, dtype=torch.float32)
        self.step_t = torch.zeros(1, device=dev, dtype=torch.float32)
        self.grad_views = []
        with torch.no_grad():
            for p, off in zip(self.params, offs):
                n = p.numel()
                view = self.flat_p[off:off + n].view(p.shape)
                view.copy_(p.detach().float())
                p.data = view
                # .grad lives in the flat gradient buffer: the hand-written backward accumulates into it directly
                # (Flow/training.py _State.out) and autograd's own accumulation adds in place, so no gather is needed
                self.grad_views.append(self.flat_g[off:off + n].view(p.shape))
                p._rfk_direct = True
        self.offsets = offs
        invalidate_caches()

    @property
    def lr(self):
        return float(self.param_groups[0]["lr"])

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = float(v)

    def state_dict(self):
        """exp_avg / exp_avg_sq / step (flat, in parameter order) + the hyper-parameters; what Solver.checkpoint stores as
        'optimizer_state_dict' (RFN/trainer.py:281)."""
        return {"state": {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                          "step": self.step_t.clone()},
                "param_groups": [{"lr": self.lr, "betas": self.betas, "eps": self.eps, "n_params": self.n_pad}]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        if g.get("n_params", self.n_pad) != self.n_pad:
            raise ValueError(f"FlatAdam.load_state_dict: flat length {g.get('n_params')} saved, {self.n_pad} here")
        self.lr, self.betas, self.eps = g["lr"], tuple(g["betas"]), g["eps"]
        with torch.no_grad():
            self.exp_avg.copy_(sd["state"]["exp_avg"])
            self.exp_avg_sq.copy_(sd["state"]["exp_avg_sq"])
            self.step_t.copy_(sd["state"]["step"])

    def zero_grad(self, set_to_none=False):
        """One memset of the flat gradient buffer; every .grad is (re)pointed at its view of it.  ``set_to_none`` is accepted
        for torch.optim compatibility but gradients stay allocated (zero), which is what makes them graph-capturable."""
        self.flat_g.zero_()
        for p, v in zip(self.params, self.grad_views):
            if p.grad is not v:
                p.grad = v

    def gather_grads(self):
        """Gradients that did not land in the flat buffer (a .grad re-assigned by user code) are copied into it; with the
        views installed by zero_grad() this is a no-op.  Absent gradients count as zero."""
        for p, v in zip(self.params, self.grad_views):
            g = p.grad
            if g is None or g is v or g.data_ptr() == v.data_ptr():
                continue
            v.copy_(g)
        return self.flat_g

    def attach(self, flow):
        """Overlap the gradient all-reduce with the backward sweep of `flow` (a ListGlow whose parameters this optimizer owns):
        the sweep runs from the prior and the deepest level up to level 1 and tells the optimizer when every gradient of a
        level is final (Flow/training.py); that level's contiguous slice of the flat gradient buffer is all-reduced
        asynchronously while the shallower -- and far more expensive -- levels are still being differentiated.  What is
        left for allreduce_grads() at the end of the step is level 1 and whatever does not belong to the flow."""
        index = {id(p): i for i, p in enumerate(self.params)}
        ranges, cur, level = {}, [], -1

        def close(key):
            ids = sorted(index[id(p)] for p in cur if id(p) in index)
            if ids and ids == list(range(ids[0], ids[-1] + 1)):
                last = self.params[ids[-1]]
                ranges[key] = (self.offsets[ids[0]], self.offsets[ids[-1]] + (last.numel() + 3) // 4 * 4)
        for m in flow.glow_frame:
            if type(m).__name__ == "Squeeze2d":
                if level >= 0:
                    close(level)
                level, cur = level + 1, []
            else:
                cur.extend(m.parameters())
        close(level)
        if getattr(flow, "learn_prior", False):
            cur = list(flow.prior.parameters())
            close("prior")
        self._ranges, self._pending, self._reduced = ranges, [], []
        flow.__dict__["_rfk_grad_hook"] = self._level_ready
        return ranges

    def _level_ready(self, key):
        if self.world <= 1 or not self.overlap or key not in getattr(self, "_ranges", {}):
            return
        import torch.distributed as dist
        lo, hi = self._ranges[key]
        self._pending.append(dist.all_reduce(self.flat_g[lo:hi], group=self.group, async_op=True))
        self._reduced.append((lo, hi))

    def allreduce_grads(self):
        """Sum over the data-parallel replicas (NCCL over NVLink); the mean's 1/world is folded into the Adam kernel.
        Slices already handed to NCCL during the backward sweep (attach()) are skipped and waited for."""
        if self.world > 1:
            import torch.distributed as dist
            done = sorted(getattr(self, "_reduced", []))
            pos = 0
            for lo, hi in done + [(self.n_pad, self.n_pad)]:
                if lo > pos:
                    dist.all_reduce(self.flat_g[pos:lo], group=self.group)
                pos = max(pos, hi)
            for w in getattr(self, "_pending", []):
                w.wait()
            self._pending, self._reduced = [], []

    def apply(self):
        self.step_t += 1.0
        call("rfk_adam_step", self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
             self.exp_avg_sq.data_ptr(), self.n_pad, self.lr, self.betas[0], self.betas[1], self.eps, 1.0 / self.world,
             self.step_t.data_ptr(), ops._stream())
        invalidate_caches()   # the kernel wrote the parameters through raw pointers
        # ... and everything derived from them that registered for it is rewritten in place: one launch per kind
        derived.REFRESHER.refresh_all(self.flat_p.device, ops._stream())

    @torch.no_grad()
    def step(self):
        self.gather_grads()
        self.allreduce_grads()
        self.apply()
