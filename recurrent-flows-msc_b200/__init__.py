"""B200-native (sm_100a) Glow flow step and ConvLSTM cell behind the reference's nn.Module interface.

``Flow`` and ``Utils`` mirror the reference's packages of the same names for the hot path only
(SURVEY.md section 8): ``from recurrent_flows_msc_b200.Flow import ListGlow`` and
``from recurrent_flows_msc_b200.Utils import ConvLSTM`` are drop-in replacements for
``from Flow import ListGlow`` / ``from Utils import ConvLSTM`` (including
``install_into`` which patches an imported reference in place).

All arithmetic runs in hand-written CUDA kernels reached through the C ABI in include/rfk.h
(librfk.so, built in-tree by ``__graft_entry__.build()``).  There is no CPU or PyTorch fallback.
"""
from . import _lib, derived, ops  # noqa: F401
from . import Flow, Utils  # noqa: F401
from .Flow import (ActNorm, AffineCoupling, Conv2dNorm, Conv2dZeros, GlowStep, InvConv, ListGlow,  # noqa: F401
                   Split2d, Squeeze2d)
from .Utils import ActFun, ConvLSTM, ConvLSTMLayer, batch_reduce, split_feature  # noqa: F401
from .parallel import shard_range, shard_batch, sync_module_state  # noqa: F401
from .graphs import Graphed, GraphedLogProb, GraphedSample, GraphedTrainStep  # noqa: F401
from .optim import FlatAdam  # noqa: F401
from .rfn_driver import time_batched_loss  # noqa: F401
from .scalers import accelerate_scalers  # noqa: F401
from .Flow.glow_modules import invalidate_caches  # noqa: F401

__version__ = "0.1.0"


def install_into(ref_flow=None, ref_utils=None):
    """Patch an already-imported reference (its ``Flow`` / ``Utils`` packages) so that models built
    afterwards (RFN, SRNN, VRNN) use the B200 kernels.  Returns the list of patched names."""
    patched = []
    if ref_flow is not None:
        for name in ("ListGlow", "ActNorm", "Conv2dZeros", "Conv2dNorm", "InvConv", "AffineCoupling",
                     "Squeeze2d", "Split2d"):
            setattr(ref_flow, name, getattr(Flow, name))
            patched.append("Flow." + name)
    if ref_utils is not None:
        for name in ("ConvLSTM", "ConvLSTMLayer"):
            setattr(ref_utils, name, getattr(Utils, name))
            patched.append("Utils." + name)
    return patched
