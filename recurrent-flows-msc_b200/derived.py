"""Batched in-place refresh of parameter-derived tensors (packed bf16 weights, per-channel affines, folded
ActNorm . InvConv matrices) after an optimizer step: ONE kernel launch per kind for everything registered, instead of one
rebuild (an rfk_pack_weight launch or ~10 ATen launches) per module (csrc/prepare.cu).

Every cache (``glow_modules._Versioned``) builds an entry the first time it is needed -- the legacy per-module path --
and registers HOW to refresh it: a row of 64-bit words (pointers of the parameters it reads and of the persistent
tensors it wrote).  ``FlatAdam.apply()`` calls ``refresh_all()`` right after the Adam kernel: the rows of each kind are
uploaded once as a device table, the batched kernels rewrite the registered outputs in place, and the cache entries are
re-stamped as valid for the new parameter epoch, so the next forward finds every cache warm and launches nothing.
Inside a captured training step the same launches are part of the optimizer graph.
"""
import weakref

import torch

from . import _lib

WORDS = 24
KINDS = ("pack", "affine", "fold")
_CALLS = {"pack": "rfk_pack_weights_batched", "affine": "rfk_affine_prepare_batched", "fold": "rfk_fold_prepare_batched"}


class _Entry:
    __slots__ = ("cache", "key", "params", "slot", "words", "keep", "elems")


class Refresher:
    def __init__(self):
        self.entries = {k: {} for k in KINDS}      # kind -> {(id(cache), key): _Entry}
        self.tables = {}                            # kind -> (device tensor, n, max_elems)
        self.dirty = {k: True for k in KINDS}

    def register(self, kind, cache, key, params, slot, words, keep, elems=0):
        e = _Entry()
        e.cache, e.key, e.params, e.slot, e.keep, e.elems = weakref.ref(cache), key, params, slot, keep, int(elems)
        e.words = [int(w) for w in words] + [0] * (WORDS - len(words))
        self.entries[kind][(id(cache), key)] = e
        self.dirty[kind] = True

    def _live(self, kind):
        """Drop entries whose cache is gone or has rebuilt the slot (re-registration replaces by key; this catches clear())."""
        dead = []
        for k, e in self.entries[kind].items():
            c = e.cache()
            if c is None or c._store.get(e.key) is not e.slot:
                dead.append(k)
        for k in dead:
            del self.entries[kind][k]
            self.dirty[kind] = True
        return list(self.entries[kind].values())

    def _table(self, kind, device):
        live = self._live(kind)
        if not live:
            self.tables.pop(kind, None)
            return None
        if self.dirty[kind] or kind not in self.tables or self.tables[kind][0].device != device:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("recurrent-flows-msc_b200: a parameter-derived tensor was first built inside a CUDA-graph capture; "
                                   "run the step eagerly once before capturing (GraphedTrainStep's warm-up does)")
            flat = [w for e in live for w in e.words]
            t = torch.tensor(flat, dtype=torch.int64).to(device)
            self.tables[kind] = (t, len(live), max(e.elems for e in live))
            self.dirty[kind] = False
        return self.tables[kind]

    def refresh_all(self, device, stream):
        """Rewrite every registered output in place from the current parameter values, then mark the entries valid."""
        for kind in ("fold", "affine", "pack"):
            tab = self._table(kind, device)
            if tab is None:
                continue
            t, n, elems = tab
            if kind == "pack":
                _lib.call(_CALLS[kind], t.data_ptr(), n, max(elems, 1), stream)
            else:
                _lib.call(_CALLS[kind], t.data_ptr(), n, stream)
        self.restamp()

    def restamp(self):
        """The registered outputs are fresh (the batched kernels ran, eagerly or in a graph replay): stamp them with the
        current parameter epoch / versions."""
        from .Flow.glow_modules import _ver_of
        for kind in KINDS:
            for e in self._live(kind):
                e.slot[0] = _ver_of(e.params)

    def clear(self):
        self.__init__()


REFRESHER = Refresher()


# ---- registration callbacks handed to _Versioned.get(..., refresh=...) ---------------------------------------------------
def reg_pack(cache, key, params, slot):
    out = slot[1][0]
    words = getattr(out, "rfk_pack", None)
    if words is not None:
        REFRESHER.register("pack", cache, key, params, slot, words, (out, getattr(out, "rfk_perm", None)), out.numel())


def reg_affine(logs, bias, factor):
    def cb(cache, key, params, slot):
        scale, shift = slot[1][0], slot[1][1]
        n = scale.numel()
        if not (logs.is_cuda and logs.dtype == torch.float32 and logs.is_contiguous()):
            return
        if bias is not None and not (bias.dtype == torch.float32 and bias.is_contiguous()):
            return
        REFRESHER.register("affine", cache, key, params, slot,
                           [logs.data_ptr(), 0 if bias is None else bias.data_ptr(), scale.data_ptr(), shift.data_ptr(), n, int(factor)],
                           (scale, shift))
    return cb
