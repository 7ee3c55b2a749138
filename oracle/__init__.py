"""CPU oracle for the Glow-step + ConvLSTM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product path (``recurrent-flows-msc_b200``) never
imports this package and fails loudly when its CUDA library is missing.

Parity status: the reference (cdglissov/recurrent-flows-msc) ships no tests or
golden vectors for this path, so the oracle is pinned
against outputs of the reference's own Python modules, generated in the build
container by ``tests/golden/make_golden.py`` (which imports /root/reference)
and committed as fixtures under ``tests/golden/``.  ``tests/test_oracle_golden.py``
replays every fixture through this oracle.
"""
from .glow_oracle import (  # noqa: F401
    squeeze2d, split_feature, batch_reduce, act_fun,
    actnorm_init, actnorm, batchnorm_flow, invconv_weight, invconv,
    conv2d_norm, conv2d_zeros, coupling_nn, clamp_log_scale,
    affine_coupling, split2d, glow_step, listglow_layout,
    listglow_f, listglow_g, listglow_prior, listglow_log_prob, listglow_sample,
    bits_per_dim,
)
from .convlstm_oracle import convlstm_cell, convlstm  # noqa: F401
