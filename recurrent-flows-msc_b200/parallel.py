"""Batch sharding for one-process-per-GPU runs.

Every op on the hot path is independent per sample (ActNorm / InvConv parameters are shared, the
log-det is per sample, ConvLSTM state is per sample), so density evaluation and sampling shard over
the batch of sequences with NO data-path collective.  The only exchange is one-off: after the
data-dependent ActNorm initialisation the parameters are broadcast from rank 0 so that all replicas
hold the same flow.  Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced [lo, hi) of `n_items` for `rank`; the first n_items % world_size ranks get one more."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world_size):
    """Slice every tensor (or list of tensors) along dim 0 to this rank's shard."""
    def one(t):
        if isinstance(t, (list, tuple)):
            return type(t)(one(u) for u in t)
        if t is None:
            return None
        lo, hi = shard_range(t.shape[0], rank, world_size)
        return t[lo:hi]
    return one(tensors)


def sync_module_state(module, src=0, group=None):
    """Broadcast parameters and buffers from `src` (e.g. after ActNorm's data-dependent init on rank 0's
    shard).  Returns the number of tensors broadcast; a no-op without an initialised process group."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    n = 0
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
            n += 1
    for m in module.modules():  # derived-tensor caches key on versions; broadcast wrote through .data
        for attr in ("_cache",):
            c = getattr(m, attr, None)
            if c is not None and hasattr(c, "clear"):
                c.clear()
        if hasattr(m, "_init_known"):
            m._init_known = None
        if hasattr(m, "_packed"):
            m._packed = None
    return n


def gather_batch(t, group=None):
    """All-gather equally-shaped per-rank results along dim 0 (evaluation bookkeeping only)."""
    if not (dist.is_available() and dist.is_initialized()):
        return t
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t.contiguous(), group=group)
    return torch.cat(out, 0)
