"""ctypes binding of librfk.so (include/rfk.h).

The library is the product; there is no fallback.  If it is missing, cannot be
loaded, or a call returns a non-zero code, an exception is raised.
"""
import ctypes
import os
from ctypes import c_float, c_int, c_longlong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librfk.so")

# name -> argtypes; every function returns int except rfk_last_error
SIGNATURES = {
    "rfk_version": [],
    "rfk_device_info": [c_void_p, c_void_p, c_void_p],
    "rfk_squeeze2d": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rfk_actnorm": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "rfk_actnorm_init": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "rfk_batch_stats_pos": [c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_float, c_void_p],
    "rfk_affine_pos": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p],
    "rfk_mix1x1": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                   c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p],
    "rfk_pack_nhwc_bf16": [c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    "rfk_copy_channels": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rfk_conv_gemm": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                      c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    "rfk_conv_gemm_coupling": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "rfk_conv_gemm_lstm": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                           c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_longlong, c_void_p, c_longlong,
                           c_void_p, c_int, c_int, c_void_p],
    "rfk_coupling_tail": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                          c_void_p],
    "rfk_coupling_tail_taps": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                               c_void_p, c_void_p, c_int, c_void_p],
    "rfk_conv1x1_taps_fused": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                               c_void_p, c_int, c_int, c_void_p, c_void_p],
    "rfk_coupling_nn_fused": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                              c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "rfk_pack_weight_folded": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_void_p],
    "rfk_coupling_taps_mix": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                              c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_float, c_void_p],
    "rfk_gauss_logp": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "rfk_gauss_sample": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_int,
                         c_void_p],
    "rfk_convlstm_pointwise": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "rfk_conv_gemm_splitk_fused": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                   c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
    "rfk_conv_gemm_splitk": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p, c_int, c_void_p],
    "rfk_convlstm_pointwise_ws": [c_void_p, c_int, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_longlong,
                                  c_void_p, c_longlong, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rfk_act_affine_bwd": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                           c_float, c_int, c_longlong, c_void_p],
    "rfk_conv_wgrad": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                       c_int, c_void_p, c_void_p, c_longlong, c_void_p],
    "rfk_coupling_taps_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                              c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p],
    "rfk_taps_scatter": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rfk_mix1x1_wgrad": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "rfk_gauss_logp_bwd": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                           c_void_p, c_void_p],
    "rfk_pack_weight": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    "rfk_adam_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_float, c_float, c_float, c_float,
                      c_void_p, c_void_p],
    "rfk_convlstm_pointwise_bwd": [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_void_p],
    "rfk_taps_gather_nhwc": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "rfk_taps_gather_nhwc_acc": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int,
                                 c_void_p],
    "rfk_add_scalar": [c_void_p, c_void_p, c_float, c_int, c_void_p],
    "rfk_set_conv_split": [c_int],
    "rfk_set_pdl": [c_int],
    "rfk_pack_nhwc_bf16_lo": [c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p],
    "rfk_pack_weights_batched": [c_void_p, c_int, c_longlong, c_void_p],
    "rfk_affine_prepare_batched": [c_void_p, c_int, c_void_p],
    "rfk_fold_prepare_batched": [c_void_p, c_int, c_void_p],
    "rfk_fold_backward_batched": [c_void_p, c_int, c_void_p, c_void_p],
    "rfk_conv_gemm_actbwd": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                             c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "rfk_actnorm_param_bwd": [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p],
    "rfk_ws_to_nchw": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "rfk_add_channels": [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "rfk_debug_set_timeline": [c_void_p, c_longlong],
}

# enums of rfk.h
ACT = {"none": 0, "relu": 1, "leakyrelu": 2}
CLAMP = {"none": 0, "realnvp": 1, "glow": 2, "softclamp": 3}
PAIR_CROSS, PAIR_SPLIT = 0, 1
STD = {"softplus": 0, "exp": 1}
OUT_NHWC_BF16, OUT_NCHW_F32 = 0, 1


class RfkError(RuntimeError):
    pass


_lib = None
launches = 0  # kernels enqueued through this binding (bench.py reports it as gpu_launches)


def lib():
    """Load librfk.so once; raise if it is absent (build with __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RfkError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = c_int
        l.rfk_last_error.argtypes = []
        l.rfk_last_error.restype = ctypes.c_char_p
        _lib = l
    return _lib


tracer = None  # optional callable(name, meta) -> context manager; bench.py uses it to time kernels with CUDA events


def call(name, *args, meta=None):
    global launches
    if tracer is not None:
        with tracer(name, meta):
            rc = getattr(lib(), name)(*args)
    else:
        rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise RfkError(f"{name} failed ({rc}): {lib().rfk_last_error().decode()}")
    launches += 1
    return rc
