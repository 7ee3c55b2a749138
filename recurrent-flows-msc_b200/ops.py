"""Tensor-level wrappers over the C ABI (include/rfk.h).

PyTorch is used here for device memory and the current stream only; every
function enqueues exactly the kernel it names on ``torch.cuda.current_stream()``.
"""
import contextlib

import torch

from . import _lib
from ._lib import ACT, CLAMP, OUT_NCHW_F32, OUT_NHWC_BF16, PAIR_CROSS, PAIR_SPLIT, STD, call  # noqa: F401


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _chk(t, dtype=torch.float32, name="tensor"):
    if not t.is_cuda:
        raise _lib.RfkError(f"{name} must be a CUDA tensor: this path has no CPU implementation")
    if t.dtype != dtype:
        raise _lib.RfkError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.RfkError(f"{name} must be contiguous")
    return t


def f32c(t):
    """Contiguous float32 CUDA view/copy of a user tensor (plumbing, not compute)."""
    if not t.is_cuda:
        raise _lib.RfkError("input must be a CUDA tensor: this path has no CPU implementation")
    return t.detach().to(torch.float32).contiguous()


def pad_to(v, m):
    return (v + m - 1) // m * m


def cin_pad(c):
    """Channels a conv consumes from an NHWC staging buffer: 32 (64-byte swizzle rows) for up to 32 real
    channels, else the next multiple of 64 (128-byte swizzle rows)."""
    return 32 if c <= 32 else pad_to(c, 64)


# Split-precision ("bf16x3") mode, RFK_CONV_PRECISION=bf16x3 (alias tf32): every conv operand is a_hi + a_lo (two bf16
# words), rows of the staging buffers are [hi | lo] halves, weights are packed [w_hi | w_hi | w_lo] per tap, and the same
# tcgen05 kernels accumulate a_hi w_hi + a_lo w_hi + a_hi w_lo in fp32 (include/rfk.h, rfk_set_conv_split).  16 significant
# bits instead of bf16's 8 (tf32: 11): the mode behind the 1e-3 parity gate.  Forward / reverse only (no training path).
import os as _os
PRECISION = {"tf32": "bf16x3", "fp32": "bf16x3"}.get(_os.environ.get("RFK_CONV_PRECISION", "bf16").lower(),
                                                    _os.environ.get("RFK_CONV_PRECISION", "bf16").lower())
if PRECISION not in ("bf16", "bf16x3"):
    raise _lib.RfkError(f"RFK_CONV_PRECISION={PRECISION!r}: expected bf16 or bf16x3 (alias tf32)")
SPLIT = PRECISION == "bf16x3"
if SPLIT:
    call("rfk_set_conv_split", 1)


@contextlib.contextmanager
def pdl(on=True):
    """Programmatic dependent launch for the launches enqueued inside the block (rfk_set_pdl)."""
    prev = _lib.lib().rfk_set_pdl(int(on))
    try:
        yield
    finally:
        _lib.lib().rfk_set_pdl(prev)


def buf_ld(c):
    """Row stride (channels) of an NHWC staging buffer for c real channels: cin_pad(c), or two such halves [hi | lo]."""
    return (2 if SPLIT else 1) * cin_pad(c)


# --------------------------------------------------------------------------------------
def squeeze2d(x, undo=False):
    _chk(x, name="x")
    B, C, H, W = x.shape
    y = torch.empty((B, C // 4, 2 * H, 2 * W) if undo else (B, 4 * C, H // 2, W // 2), device=x.device, dtype=x.dtype)
    call("rfk_squeeze2d", x.data_ptr(), y.data_ptr(), B, C, H, W, int(undo), _stream(), meta={"bytes": 8.0 * x.numel()})
    return y


def actnorm(x, bias, logs, reverse=False):
    _chk(x, name="x")
    B, C, H, W = x.shape
    y = torch.empty_like(x)
    call("rfk_actnorm", x.data_ptr(), y.data_ptr(), _chk(bias).data_ptr(), _chk(logs).data_ptr(), B, C, H * W,
         int(reverse), _stream(), meta={"bytes": 8.0 * x.numel()})
    return y


def actnorm_init(x, bias, logs):
    """Writes bias/logs in place from the statistics of x [B,C,H,W]."""
    _chk(x, name="x")
    B, C, H, W = x.shape
    call("rfk_actnorm_init", x.data_ptr(), _chk(bias).data_ptr(), _chk(logs).data_ptr(), 0, 0, B, C, H * W, _stream())


def batch_stats_pos(x, eps):
    """Per-position (C,H,W) mean and biased variance + eps over the batch; returns ([C,H,W], [C,H,W])."""
    _chk(x, name="x")
    B, C, H, W = x.shape
    mean = torch.empty(C, H, W, device=x.device, dtype=torch.float32)
    var = torch.empty_like(mean)
    call("rfk_batch_stats_pos", x.data_ptr(), mean.data_ptr(), var.data_ptr(), B, C * H * W, float(eps), _stream())
    return mean, var


def affine_pos(x, a, c):
    """y[b,i] = x[b,i]*a[i] + c[i] with a, c of shape [C,H,W] (contiguous f32)."""
    _chk(x, name="x")
    B = x.shape[0]
    y = torch.empty_like(x)
    call("rfk_affine_pos", x.data_ptr(), y.data_ptr(), _chk(a).data_ptr(), _chk(c).data_ptr(), B, x[0].numel(), _stream())
    return y


def channel_stats(x, mean, std):
    """Per-channel mean and UNBIASED std of x [B,C,H,W] over (B,H,W) into the given [C] tensors."""
    _chk(x, name="x")
    B, C, H, W = x.shape
    call("rfk_actnorm_init", x.data_ptr(), 0, 0, _chk(mean).data_ptr(), _chk(std).data_ptr(), B, C, H * W, _stream())


def mix1x1(x, Wm, bvec=None, side=None, side_n=0, side_off=0, logdet=None, addend=None, alpha=1.0):
    """y = Wm x + bvec per pixel; optionally logdet[b] += alpha * addend (device scalar) in the same launch."""
    _chk(x, name="x")
    B, C, H, W = x.shape
    y = torch.empty_like(x)
    if SPLIT and side is not None:    # the side output is a single bf16 word per value: pack hi and lo from y instead
        call("rfk_mix1x1", x.data_ptr(), y.data_ptr(), _chk(Wm).data_ptr(), _p(bvec), B, C, H * W, 0, 0, 0, 0, _p(logdet),
             _p(addend), float(alpha), _stream(), meta={"bytes": 8.0 * x.numel()})
        pack_nhwc(y, 0, side_n, side, side_off)
        return y
    side_ld = side.shape[-1] if side is not None else 0
    call("rfk_mix1x1", x.data_ptr(), y.data_ptr(), _chk(Wm).data_ptr(), _p(bvec), B, C, H * W,
         _p(side), side_n, side_off, side_ld, _p(logdet), _p(addend), float(alpha), _stream(),
         meta={"bytes": 8.0 * x.numel() + 2.0 * B * H * W * (side_n if side is not None else 0)})   # x in, y out, bf16 side out
    return y


def pack_nhwc(src, c_lo, n, dst, dst_off):
    """src NCHW f32 channels [c_lo,c_lo+n) -> dst NHWC bf16 [B,H,W,ld] at channel dst_off.
    src may be a batch-strided slice (e.g. x[:, t] of a [B,T,C,H,W] sequence)."""
    B, C, H, W = src.shape
    if n == 0:
        return
    if not src.is_cuda or src.dtype != torch.float32 or src[0].is_contiguous() is False:
        raise _lib.RfkError("src must be a float32 CUDA tensor, dense within each sample")
    _chk(dst, torch.bfloat16, "dst")
    call("rfk_pack_nhwc_bf16", src.data_ptr(), src.stride(0), B, C, H * W, c_lo, n, dst.data_ptr(), dst_off,
         dst.shape[-1], _stream(), meta={"bytes": 6.0 * B * n * H * W})
    if SPLIT:
        call("rfk_pack_nhwc_bf16_lo", src.data_ptr(), src.stride(0), B, C, H * W, c_lo, n, dst.data_ptr(),
             dst_off + dst.shape[-1] // 2, dst.shape[-1], _stream(), meta={"bytes": 6.0 * B * n * H * W})


def copy_channels(src, src_off, dst, dst_off, n):
    _chk(src, name="src")
    _chk(dst, name="dst")
    B, Cs, H, W = src.shape
    call("rfk_copy_channels", src.data_ptr(), Cs, src_off, dst.data_ptr(), dst.shape[1], dst_off, n, B, H * W, _stream())


def _gemm_meta(M, n, taps, cin_pad, wgt):
    """Algorithmic (unpadded) and issued (padded) FLOPs of one implicit-GEMM launch, for bench.py's roofline."""
    cin = getattr(wgt, "rfk_cin", cin_pad)
    if SPLIT:
        cin, cin_pad = 3 * cin, 3 * cin_pad    # three bf16 products per real multiply
    return {"flops": 2.0 * M * n * taps * cin, "flops_padded": 2.0 * M * wgt.shape[0] * taps * cin_pad,
            "M": M, "N": n, "K": taps * cin,
            # algorithmic HBM bytes: activations read once (bf16, padded row), output written once as bf16
            # (f32 outputs of the small-N launches are counted as 4 B), weights once
            "bytes": 2.0 * M * cin_pad + (2.0 if n >= 128 else 4.0) * M * n + 2.0 * wgt.numel()}


def conv_gemm(act, cin_pad, wgt, n, taps, scale, shift, act_fn, out, out_off=0):
    """act NHWC bf16 [B,H,W,ld]; wgt bf16 [n_pad, taps*cin_pad]; out NHWC bf16 (4-D, channel last) or NCHW f32."""
    _chk(act, torch.bfloat16, "act")
    _chk(wgt, torch.bfloat16, "wgt")
    B, H, W, ld = act.shape
    if out.dtype == torch.bfloat16:
        kind, out_ld = OUT_NHWC_BF16, out.shape[-1]
        _chk(out, torch.bfloat16, "out")
    else:
        kind, out_ld = OUT_NCHW_F32, 0
        _chk(out, name="out")
    call("rfk_conv_gemm", act.data_ptr(), B, H, W, ld, cin_pad, wgt.data_ptr(), n, wgt.shape[0], taps,
         _p(scale), _p(shift), ACT[act_fn], kind, out.data_ptr(), out_ld, out_off, _stream(),
         meta=_gemm_meta(B * H * W, n, taps, cin_pad, wgt))
    return out


def conv_gemm_coupling(act, cin_pad, wgt, n, taps, scale, shift, z, clamp_type, clamp_scale, clamp_shift,
                       logdet, reverse):
    _chk(act, torch.bfloat16, "act")
    _chk(z, name="z")
    B, H, W, ld = act.shape
    call("rfk_conv_gemm_coupling", act.data_ptr(), B, H, W, ld, cin_pad, _chk(wgt, torch.bfloat16).data_ptr(), n,
         wgt.shape[0], taps, _p(scale), _p(shift), z.data_ptr(), CLAMP[clamp_type], _p(clamp_scale), _p(clamp_shift),
         _p(logdet), int(reverse), _stream(), meta=_gemm_meta(B * H * W, n, taps, cin_pad, wgt))


def conv_gemm_lstm(act, cin_pad, wgt, hidden, ht, ht_pad, taps, bias, c_prev, peep, c_next, h_out, h_nhwc, h_off):
    _chk(act, torch.bfloat16, "act")
    B, H, W, ld = act.shape
    call("rfk_conv_gemm_lstm", act.data_ptr(), B, H, W, ld, cin_pad, _chk(wgt, torch.bfloat16).data_ptr(), hidden, ht,
         ht_pad, taps, _p(bias), _p(c_prev), c_prev.stride(0) if c_prev is not None else 0, _p(peep),
         c_next.data_ptr(), c_next.stride(0), h_out.data_ptr(), h_out.stride(0),
         _p(h_nhwc), h_off, h_nhwc.shape[-1] if h_nhwc is not None else 0, _stream(),
         meta=_gemm_meta(B * H * W, 4 * hidden, taps, cin_pad, wgt))


def choose_k_split(M, taps, cin_pad):
    """K slices for a 3x3 conv whose pixel tiles cannot fill the GPU.  tcgen05.mma (M=128) costs ~131 cycles whatever
    N is, so a CTA's time is K-iterations x 4 x 131 cycles however narrow its tile: the only way to put more SMs on a
    small-M layer is to cut K (one slice per filter tap)."""
    # Measured on B200 (tools/splitk_bench.py, 30 sequences): with per-slice slabs in a tile-local layout and a cooperative
    # fix-up the split kernel's main loop takes 4 us but the hand-shake + slab reduction add ~12 us, so it wins only on
    # the deepest level (one pixel tile, K = 9*320: 24.1 -> 16.4 us) and loses on 4x4 / 8x8 maps (17.3 vs 16.3, 17.9 vs
    # 12.7 us).  RFK_CONV_SPLITK=0 disables it, =2 also splits up to SPLITK_MAX_UNITS (pixel tiles x slices).
    import os
    mode = os.environ.get("RFK_CONV_SPLITK", "1")
    if mode == "0" or SPLIT or taps != 9 or cin_pad % 64 != 0 or cin_pad < 128:
        return 1
    m_tiles = (M + 127) // 128
    if m_tiles == 1 and taps * cin_pad >= 2048:
        return 9
    if 16 <= m_tiles and m_tiles * 3 <= SPLITK_MAX_UNITS and taps * cin_pad >= 2048:
        # the deepest level of the 570-frame workload (18 pixel tiles, K = 9*320): three filter rows per pixel tile put 54 SMs
        # on it instead of 18 (training step 26.8 -> 26.4 ms on the same box)
        return 3
    if mode == "2":
        if m_tiles * 9 <= SPLITK_MAX_UNITS:
            return 9
        if m_tiles * 3 <= SPLITK_MAX_UNITS:
            return 3
    return 1


SPLITK_MAX_UNITS = 74


def gemm_m_tiles(B, H, W):
    """Number of 128-pixel tiles the conv kernel cuts [B,H,W] into (tile = NIMG x TH x TW with power-of-two TW, TH)."""
    twl = min(max(W - 1, 0).bit_length(), 7)
    thl = min(max(H - 1, 0).bit_length(), 7 - twl)
    TW, TH = 1 << twl, 1 << thl
    nimg = 128 // (TW * TH)
    return -(-W // TW) * -(-H // TH) * -(-B // nimg)


def conv_gemm_splitk_fused(act, cin_pad, wgt, n, taps, k_split, scale, shift, act_fn, out, out_off=0):
    """conv_gemm (bf16 NHWC output) with K cut into k_split slices and the reduction fused into the kernel."""
    _chk(act, torch.bfloat16, "act")
    _chk(out, torch.bfloat16, "out")
    B, H, W, ld = act.shape
    n_pad = wgt.shape[0]
    m_tiles = gemm_m_tiles(B, H, W)
    ws = workspace(("splitk_ws", n_pad, k_split, m_tiles), (k_split * m_tiles * 128, n_pad), act.device, torch.float32)   # one slab per slice
    cnt = workspace(("splitk_cnt",), (max(1024, 4 * m_tiles),), act.device, torch.int32)       # zero, kept zero
    call("rfk_conv_gemm_splitk_fused", act.data_ptr(), B, H, W, ld, cin_pad, _chk(wgt, torch.bfloat16).data_ptr(), n,
         n_pad, taps, k_split, ws.data_ptr(), n_pad, cnt.data_ptr(), _p(scale), _p(shift), ACT[act_fn],
         out.data_ptr(), out.shape[-1], out_off, _stream(), meta=_gemm_meta(B * H * W, n, taps, cin_pad, wgt))
    return out


def conv_gemm_splitk(act, cin_pad, wgt, n, taps, k_split, ws):
    """Partial sums of k_split K slices added into the zeroed fp32 workspace ws [B*H*W, ld]."""
    _chk(act, torch.bfloat16, "act")
    _chk(ws, name="ws")
    B, H, W, ld = act.shape
    call("rfk_conv_gemm_splitk", act.data_ptr(), B, H, W, ld, cin_pad, _chk(wgt, torch.bfloat16).data_ptr(), n,
         wgt.shape[0], taps, k_split, ws.data_ptr(), ws.shape[-1], _stream(),
         meta=_gemm_meta(B * H * W, n, taps, cin_pad, wgt))


def conv_gemm_small(act, cin_pad, wgt, n, taps, bias, out):
    """conv (fp32 NCHW out, + bias) for launches of one or two pixel tiles with a long K (3x3, >= 256 channels: the ConvLSTM
    of RFN on 2x2 maps): K is cut into one slice per filter tap so that nine times as many SMs stream the weights (a single
    CTA is limited by its own L2 bandwidth), the slices add into a zero fp32 accumulator, and a tiny kernel moves the sum
    into NCHW and leaves the accumulator zero again."""
    B, H, W, _ = act.shape
    ws = workspace(("small_ws", wgt.shape[0]), (B * H * W, wgt.shape[0]), act.device, torch.float32)   # zero between uses
    conv_gemm_splitk(act, cin_pad, wgt, n, taps, taps, ws)
    call("rfk_ws_to_nchw", ws.data_ptr(), ws.shape[-1], _p(bias), _chk(out, name="out").data_ptr(), B, n, H * W, 1, _stream())
    return out


def use_small_gemm(B, H, W, taps, cin_pad):
    return taps == 9 and B * H * W <= 256 and cin_pad >= 256 and not SPLIT


def convlstm_pointwise_ws(cc, bias, c_prev, peep, h_out, c_next, h_nhwc, h_off, zero_cc=True):
    B, Hc, H, W = c_next.shape
    call("rfk_convlstm_pointwise_ws", _chk(cc).data_ptr(), cc.shape[-1], _p(bias), _p(c_prev),
         c_prev.stride(0) if c_prev is not None else 0, _p(peep), h_out.data_ptr(), h_out.stride(0), c_next.data_ptr(),
         c_next.stride(0), _p(h_nhwc), h_off, h_nhwc.shape[-1] if h_nhwc is not None else 0, B, Hc, H * W,
         int(zero_cc), _stream(),
         meta={"bytes": 4.0 * B * Hc * H * W * ((8 if zero_cc else 4) + (1 if c_prev is not None else 0) + 2) +
                        (2.0 * B * Hc * H * W if h_nhwc is not None else 0)})   # gates in (+cleared), c in, h and c out, bf16 h


def coupling_tail(nn_out, z, clamp_type, clamp_scale, clamp_shift, logdet, reverse):
    _chk(nn_out, name="nn_out")
    _chk(z, name="z")
    B, C, H, W = z.shape
    call("rfk_coupling_tail", nn_out.data_ptr(), z.data_ptr(), B, C, H * W, CLAMP[clamp_type], _p(clamp_scale),
         _p(clamp_shift), _p(logdet), int(reverse), _stream())


def coupling_tail_taps(taps, z, scale, shift, clamp_type, clamp_scale, clamp_shift, logdet, reverse):
    _chk(taps, name="taps")
    _chk(z, name="z")
    B, C, H, W = z.shape
    call("rfk_coupling_tail_taps", taps.data_ptr(), z.data_ptr(), B, C, H, W, _chk(scale).data_ptr(),
         _chk(shift).data_ptr(), CLAMP[clamp_type], _p(clamp_scale), _p(clamp_shift), _p(logdet), int(reverse), _stream(),
         meta={"bytes": 4.0 * taps.numel() + 4.0 * z.numel()})   # nine tap planes in, z2 half read + written


def conv1x1_taps_fused(act, cin_pad, w2, hid, scale2, shift2, act_fn, w9, n3, taps):
    """taps = tap-split conv3x3( act_fn(scale2 * conv1x1(act) + shift2) ) with the hidden tensor kept in tensor memory."""
    _chk(act, torch.bfloat16, "act")
    _chk(taps, name="taps")
    B, H, W, ld = act.shape
    M = B * H * W
    meta = {"flops": 2.0 * M * hid * cin_pad + 2.0 * M * n3 * hid,
            "flops_padded": 2.0 * M * hid * cin_pad + 2.0 * M * w9.shape[0] * hid,
            "M": M, "N": hid, "K": cin_pad, "bytes": 2.0 * M * cin_pad + 4.0 * M * n3 + 2.0 * (w2.numel() + w9.numel())}
    call("rfk_conv1x1_taps_fused", act.data_ptr(), B, H, W, ld, cin_pad, _chk(w2, torch.bfloat16).data_ptr(), hid,
         _p(scale2), _p(shift2), ACT[act_fn], _chk(w9, torch.bfloat16).data_ptr(), n3, w9.shape[0], taps.data_ptr(),
         _stream(), meta=meta)
    return taps


def coupling_nn_fused(act, cin_pad, taps_k, w1f, hid, w2f, act_fn, w9, n3, taps, h1=None, h2=None):
    """The whole coupling network in one kernel: taps = tap-split conv3x3( act(ActNorm(conv1x1( act(ActNorm(conv(act))) ))) ),
    both hidden tensors in tensor memory.  w1f / w2f: weights with their ActNorm folded in (pack_conv_weight_folded);
    h1 / h2 (NHWC bf16, optional): side outputs for the backward pass."""
    _chk(act, torch.bfloat16, "act")
    _chk(taps, name="taps")
    B, H, W, ld = act.shape
    M = B * H * W
    cin = getattr(w1f, "rfk_cin", cin_pad)
    store = h1 is not None
    assert w1f.shape == (hid, taps_k * cin_pad + 16) and w2f.shape == (hid, hid + 16), "folded weights: 16 shift columns behind K"
    meta = {"flops": 2.0 * M * hid * (taps_k * cin + hid + n3),
            "flops_padded": 2.0 * M * hid * (taps_k * cin_pad + 16 + hid + 16 + pad_to(n3, 16)),
            "M": M, "N": hid, "K": taps_k * cin,
            "bytes": 2.0 * M * cin_pad + 4.0 * M * n3 + 2.0 * (w1f.numel() + w2f.numel() + w9.numel())
                     + (4.0 * M * hid if store else 0.0)}
    call("rfk_coupling_nn_fused", act.data_ptr(), B, H, W, ld, cin_pad, taps_k, _chk(w1f, torch.bfloat16).data_ptr(), hid,
         _chk(w2f, torch.bfloat16).data_ptr(), ACT[act_fn], _chk(w9, torch.bfloat16).data_ptr(), n3, w9.shape[0],
         taps.data_ptr(), _chk(h1, torch.bfloat16).data_ptr() if store else None,
         _chk(h2, torch.bfloat16).data_ptr() if store else None, h1.shape[-1] if store else 0, _stream(), meta=meta)
    return taps


def coupling_taps_mix(taps, z, scale, shift, clamp_type, clamp_scale, clamp_shift, cpl_logdet, reverse, Wm, bvec,
                      side=None, side_n=0, side_off=0, logdet=None, addend=None, alpha=1.0):
    """y = Wm * coupling(z; taps) + bvec: the tap gather + coupling update of one step fused with the next 1x1 mix."""
    _chk(taps, name="taps")
    _chk(z, name="z")
    B, C, H, W = z.shape
    y = torch.empty_like(z)
    side_ld = side.shape[-1] if side is not None else 0
    call("rfk_coupling_taps_mix", taps.data_ptr(), z.data_ptr(), y.data_ptr(), B, C, H, W, _chk(scale).data_ptr(),
         _chk(shift).data_ptr(), CLAMP[clamp_type], _p(clamp_scale), _p(clamp_shift), _p(cpl_logdet), int(reverse),
         _chk(Wm).data_ptr(), _p(bvec), _p(side), side_n, side_off, side_ld, _p(logdet), _p(addend), float(alpha), _stream())
    return y


def gauss_logp(z, z_off, params, n, pairing, std_kind, logdet):
    _chk(z, name="z")
    B, zC, H, W = z.shape
    call("rfk_gauss_logp", z.data_ptr(), zC, z_off, _p(params), n, B, H * W, pairing, STD[std_kind],
         _chk(logdet).data_ptr(), _stream(), meta={"bytes": 4.0 * B * n * H * W * (3 if params is not None else 1)})


def gauss_sample(eps, params, n, pairing, std_kind, temperature, out, out_off):
    _chk(eps, name="eps")
    _chk(out, name="out")
    B, oC, H, W = out.shape
    call("rfk_gauss_sample", eps.data_ptr(), _p(params), n, B, H * W, pairing, STD[std_kind], float(temperature),
         out.data_ptr(), oC, out_off, _stream())


def add_channels(dst, dst_off, src, src_off, n):
    """dst[:, dst_off:dst_off+n] += src[:, src_off:src_off+n] (fp32 NCHW, same B, H, W)."""
    _chk(dst, name="dst")
    _chk(src, name="src")
    B, Cd, H, W = dst.shape
    call("rfk_add_channels", dst.data_ptr(), Cd, dst_off, src.data_ptr(), src.shape[1], src_off, n, B, H * W, _stream())


def convlstm_pointwise(cc, c_prev, peep):
    _chk(cc, name="cc")
    _chk(c_prev, name="c_prev")
    B, Hc, H, W = c_prev.shape
    h = torch.empty_like(c_prev)
    c = torch.empty_like(c_prev)
    call("rfk_convlstm_pointwise", cc.data_ptr(), c_prev.data_ptr(), _p(peep), h.data_ptr(), c.data_ptr(), B, Hc,
         H * W, _stream())
    return h, c


# --------------------------------------------------------------------------------------
# backward building blocks
# --------------------------------------------------------------------------------------
class _ZeroArena:
    """One memset for all the small zero-initialised reduction / accumulation buffers of a backward sweep (per-channel
    sums, weight gradients accumulated with atomics) instead of one fill kernel each (~450 per training step)."""

    CHUNK = 1 << 23   # floats

    def __init__(self, device):
        self.device, self.buf, self.off = device, None, 0

    def take(self, n):
        n4 = (n + 3) // 4 * 4
        if self.buf is None or self.off + n4 > self.buf.numel():
            self.buf = torch.zeros(max(self.CHUNK, n4), device=self.device, dtype=torch.float32)
            self.off = 0
        out = self.buf[self.off:self.off + n]
        self.off += n4
        return out


_ARENA = None


@contextlib.contextmanager
def zero_arena(device):
    global _ARENA
    prev, _ARENA = _ARENA, _ZeroArena(device)
    try:
        yield
    finally:
        _ARENA = prev


def _zeros(n, device):
    if _ARENA is not None and _ARENA.device == device:
        return _ARENA.take(n)
    return torch.zeros(n, device=device, dtype=torch.float32)


def act_affine_bwd(dh, h, n, scale, act_fn, dvv_factor=1.0, dv_scaled=False, out_dv=None, out_dvv=None):
    """Backward of h = act(a*scale + shift): returns (da bf16 NHWC like dh, r_dv [n], r_dvv [n]) with r_dv = sum dv
    (times scale when dv_scaled: d bias of an ActNorm with scale = exp(logs)) and r_dvv = dvv_factor * sum dv*v (d logs).
    ``out_dv`` / ``out_dvv``: fp32 buffers of n elements the reductions are ACCUMULATED into (e.g. the parameters' .grad)."""
    _chk(dh, torch.bfloat16, "dh")
    _chk(h, torch.bfloat16, "h")
    rows = dh.numel() // dh.shape[-1]
    da = torch.empty_like(dh) if (n + 7) // 8 * 8 == dh.shape[-1] else torch.zeros_like(dh)   # the kernel writes 8-channel groups
    if out_dv is None or out_dvv is None:
        r = _zeros(2 * n, dh.device)
        out_dv, out_dvv = r[:n], r[n:]
    call("rfk_act_affine_bwd", dh.data_ptr(), h.data_ptr(), dh.shape[-1], n, _chk(scale).data_ptr(), ACT[act_fn],
         da.data_ptr(), da.shape[-1], out_dv.data_ptr(), out_dvv.data_ptr(), float(dvv_factor), int(bool(dv_scaled)), rows,
         _stream())
    return da, out_dv, out_dvv


def conv_gemm_actbwd(act, cin_pad, wgt, n, taps, scale, act_fn, h, out, colsum):
    """out (bf16 NHWC) = conv(act, wgt) * act'(h) * scale: a data gradient fused with the backward of the activation +
    ActNorm that produced the conv's input h; colsum [n] += column sums of out (include/rfk.h, rfk_conv_gemm_actbwd)."""
    _chk(act, torch.bfloat16, "act")
    _chk(h, torch.bfloat16, "h")
    _chk(out, torch.bfloat16, "out")
    B, H, W, ld = act.shape
    meta = _gemm_meta(B * H * W, n, taps, cin_pad, wgt)
    meta["bytes"] += 2.0 * B * H * W * n      # + the saved activation read by the epilogue
    call("rfk_conv_gemm_actbwd", act.data_ptr(), B, H, W, ld, cin_pad, _chk(wgt, torch.bfloat16).data_ptr(), n, wgt.shape[0],
         taps, _p(scale), ACT[act_fn], h.data_ptr(), h.shape[-1], out.data_ptr(), out.shape[-1],
         _chk(colsum).data_ptr(), _stream(), meta=meta)
    return out


def actnorm_param_bwd(weight, dW, colsum, bias, grad_W, d_logs, d_bias):
    """ActNorm gradients of a Conv2dNorm from its own weight gradient dW and the column sums of da; grad_W += dW."""
    n = weight.shape[0]
    call("rfk_actnorm_param_bwd", _chk(weight).data_ptr(), _chk(dW).data_ptr(), weight[0].numel(), _chk(colsum).data_ptr(),
         _chk(bias).data_ptr(), _chk(grad_W).data_ptr(), _chk(d_logs).data_ptr(), _chk(d_bias).data_ptr(), n, _stream())


def convlstm_pointwise_bwd(cc, c_prev, peep, dh, dc_in, dbias):
    """Backward of the ConvLSTM cell update: returns (dcc [B,4Hc,H,W], dc_prev [B,Hc,H,W]); dbias [4Hc] accumulates."""
    _chk(cc, name="cc")
    B, Hc4, H, W = cc.shape
    Hc = Hc4 // 4
    if not dh.is_cuda or dh.dtype != torch.float32 or not dh[0].is_contiguous():
        raise _lib.RfkError("dh must be a float32 CUDA tensor, dense within each sample")
    dcc = torch.empty_like(cc)
    dc_prev = torch.empty(B, Hc, H, W, device=cc.device, dtype=torch.float32)
    call("rfk_convlstm_pointwise_bwd", cc.data_ptr(), _p(c_prev), _p(peep), dh.data_ptr(), dh.stride(0), _p(dc_in),
         dcc.data_ptr(), dc_prev.data_ptr(), _p(dbias), B, Hc, H * W, _stream())
    return dcc, dc_prev


_WGRAD_WS = {}


def _wgrad_ws(device, slot=0):
    """Scratch for the weight-gradient kernel's per-slice partial tiles: 64 MB per slot.  Launches that share a slot must be
    stream-ordered (slot 0: the main stream; slot 1: the side stream of Flow/training.py)."""
    hit = _WGRAD_WS.get((device, slot))
    if hit is None:
        hit = _WGRAD_WS[(device, slot)] = torch.empty(1 << 24, device=device, dtype=torch.float32)
    return hit


def conv_wgrad(x, cin, dy, cout, taps, out=None, perm=None, ws_slot=0, dw_ld=None):
    """Weight gradient [cout, cin, k, k] (fp32, the conv weight's own layout) of a 'same' conv from NHWC bf16 activations x
    and output gradients dy.  ``perm`` (long tensor): channel c of x is the weight's input channel perm[c].  ``out``
    (zeroed by the caller once) accumulates over several calls."""
    _chk(x, torch.bfloat16, "x")
    _chk(dy, torch.bfloat16, "dy")
    B, H, W, xld = x.shape
    k = 3 if taps == 9 else 1
    dw = out if out is not None else _zeros(cout * cin * taps, x.device).view(cout, cin, k, k)
    _chk(dw, name="dw")
    ws = _wgrad_ws(x.device, ws_slot)
    call("rfk_conv_wgrad", x.data_ptr(), xld, cin, dy.data_ptr(), dy.shape[-1], cout, B, H, W, taps, dw.data_ptr(),
         cin if dw_ld is None else dw_ld, 1, _p(_perm32(perm)), ws.data_ptr(), ws.numel() * 4, _stream(),
         meta={"flops": 2.0 * B * H * W * cout * cin * taps, "flops_padded": 2.0 * B * H * W * cout * cin * taps,
               "M": B * H * W, "N": cout, "K": taps * cin, "bytes": 2.0 * B * H * W * (cin + cout)})
    return dw


def coupling_taps_bwd(taps, z_out, dz, scale, shift, clamp, clamp_scale, clamp_shift, g_ld, logs_factor=0.0, outs=None):
    """Backward of coupling_tail_taps.  dz (in/out, [B,C,H,W]): z2 half replaced by the gradient w.r.t. z2.
    Returns (dsum [B,C,H,W], d_scale [C], d_shift [C], d_clamp_scale [C/2], d_clamp_shift [C/2]); with logs_factor f != 0
    the second and third are d logs and d bias of a Conv2dZeros whose affine is scale = exp(f*logs), shift = bias*scale."""
    B, C, H, W = _chk(dz, name="dz").shape
    _chk(taps, name="taps"); _chk(z_out, name="z_out")
    dsum = torch.empty_like(dz)
    if outs is None:     # else: four fp32 buffers ([C], [C], [C/2], [C/2]) the reductions are ACCUMULATED into
        r = _zeros(3 * C, dz.device)
        outs = r[:C], r[C:2 * C], r[2 * C:2 * C + C // 2], r[2 * C + C // 2:]
    d_scale, d_shift, d_cs, d_csh = outs
    p = lambda t: _chk(t).data_ptr() if t is not None else None
    call("rfk_coupling_taps_bwd", taps.data_ptr(), z_out.data_ptr(), dz.data_ptr(), dsum.data_ptr(), B, C, H, W,
         _chk(scale).data_ptr(), _chk(shift).data_ptr(), CLAMP[clamp], p(clamp_scale), p(clamp_shift), p(g_ld),
         d_scale.data_ptr(), d_shift.data_ptr(), d_cs.data_ptr(), d_csh.data_ptr(), float(logs_factor), _stream())
    return dsum, d_scale, d_shift, d_cs, d_csh


def taps_scatter(dsum, out=None):
    """NHWC bf16 gradient of the nine tap planes [B,H,W,pad64(9C)] from dsum [B,C,H,W] (pad columns zero)."""
    B, C, H, W = _chk(dsum, name="dsum").shape
    ld = pad_to(9 * C, 64)
    if out is None:
        out = torch.zeros(B, H, W, ld, device=dsum.device, dtype=torch.bfloat16)
    call("rfk_taps_scatter", dsum.data_ptr(), out.data_ptr(), ld, B, C, H, W, _stream())
    return out


def mix1x1_wgrad(x, dy, out=None):
    """(dW [C,C], db [C]) of y = W x + b over all pixels (fp32 NCHW); ``out``: zeroed fp32 buffer of C*C + C elements."""
    B, C, H, W = _chk(x, name="x").shape
    _chk(dy, name="dy")
    r = _zeros(C * C + C, x.device) if out is None else out
    call("rfk_mix1x1_wgrad", x.data_ptr(), dy.data_ptr(), B, C, H * W, r.data_ptr(), r[C * C:].data_ptr(), _stream())
    return r[:C * C].view(C, C), r[C * C:C * C + C]


def gauss_logp_bwd(z, z_off, n, params, pairing, std_kind, g, dz):
    """Backward of gauss_logp: dz[:, z_off:z_off+n] += dlogp/dz * g[b]; returns dparams [B,2n,H,W] (None without params)."""
    B, zC, H, W = _chk(z, name="z").shape
    _chk(dz, name="dz"); _chk(g, name="g")
    dparams = torch.empty_like(params) if params is not None else None
    call("rfk_gauss_logp_bwd", z.data_ptr(), zC, z_off, _p(params), n, B, H * W, pairing, STD[std_kind], g.data_ptr(), dz.data_ptr(), dparams.data_ptr() if dparams is not None else None,
         _stream())
    return dparams


def pack_dgrad_taps_weight(weight, out_perm=None):
    """Tap-split form of the data-gradient weights of a 3x3 conv: a 1x1 weight [pad16(9*R8), cin_pad(N)] whose row
    t*R8 + j is W[:, perm[j], 8-t] (R8 = rows per tap rounded up to 8, the extra rows zero); the nine planes of the GEMM
    output are summed with their shifts by taps_gather_nhwc.  Returns (weight, cin_pad, R8)."""
    N, Cin, kh, kw = weight.shape
    R = Cin if out_perm is None else out_perm.numel()
    R8 = pad_to(R, 8)
    perm = None
    if out_perm is not None:
        perm = torch.full((R8,), -1, device=out_perm.device, dtype=out_perm.dtype)
        perm[:R] = out_perm
    kp = cin_pad(N)
    wgt, _ = _pack_weight(weight, 3, perm, kh * kw * R8, pad_to(kh * kw * R8, 64), kp, kp, N)   # 64: a wide N tile always divides it
    return wgt, kp, R8


def taps_gather_nhwc(T, n, n_stride, out):
    """out [B,n,H,W] fp32 = sum over the nine shifted planes of T (NHWC bf16, channel t*n_stride + j)."""
    _chk(T, torch.bfloat16, "T")
    _chk(out, name="out")
    B, H, W, ld = T.shape
    call("rfk_taps_gather_nhwc", T.data_ptr(), ld, n, n_stride, B, H, W, out.data_ptr(), _stream())
    return out


def taps_gather_nhwc_acc(T, n, n_stride, acc0, n0, acc1):
    """The gather of taps_gather_nhwc accumulated in place: channels [0, n0) += into acc0[:, :n0], channels [n0, n) into
    acc1[:, :n - n0] (fp32 NCHW)."""
    _chk(T, torch.bfloat16, "T")
    B, H, W, ld = T.shape
    if n0 > 0:
        _chk(acc0, name="acc0")
    if n0 < n:
        _chk(acc1, name="acc1")
    call("rfk_taps_gather_nhwc_acc", T.data_ptr(), ld, n, n_stride, B, H, W, acc0.data_ptr() if n0 > 0 else None,
         acc0.shape[1] if n0 > 0 else 0, n0, acc1.data_ptr() if n0 < n else None, acc1.shape[1] if n0 < n else 0, _stream(),
         meta={"bytes": 2.0 * B * H * W * 9 * n_stride + 8.0 * B * H * W * n})
    return acc1


_PERM32 = {}


def _perm32(perm):
    """int32 device copy of a (cached, long) channel permutation, for the packing kernel."""
    if perm is None:
        return None
    key = (perm.data_ptr(), perm.numel(), perm.device)
    hit = _PERM32.get(key)
    if hit is None:
        hit = _PERM32[key] = (perm, perm.to(torch.int32).contiguous())   # keeps `perm` alive: the key is its address
    return hit[1]


def _split3(packed_f32, taps, kp):
    """[rows, taps*kp] fp32 in GEMM layout -> bf16 [rows, taps*3*kp] with [w_hi | w_hi | w_lo] per tap."""
    rows = packed_f32.shape[0]
    w = packed_f32.view(rows, taps, kp)
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, hi, lo], 2).reshape(rows, taps * 3 * kp).contiguous()


def _pack_weight_split(weight, mode, perm, rows, rows_pad, kp, ktot, cin_real):
    """Split-precision form of _pack_weight: the same index maps, built with torch ops (cached per parameter version)."""
    w = weight.detach().float()
    N, Cin, kh, kw = w.shape
    taps = kh * kw
    w = w.reshape(N, Cin, taps)
    dev = w.device
    if mode == 0:      # row n, k = t*kp + j <- W[n, perm[j], t]
        src = w if perm is None else w[:, perm.long().clamp_min(0)] * (perm >= 0).float()[None, :, None]
        full = torch.zeros(rows_pad, taps, kp, device=dev)
        full[:N, :, :src.shape[1]] = src.permute(0, 2, 1)
        k_taps = taps
    elif mode == 1:    # row r, k = t*kp + co <- W[co, perm[r], taps-1-t]
        sel = w if perm is None else w[:, perm.long()]
        full = torch.zeros(rows_pad, taps, kp, device=dev)
        full[:rows, :, :N] = sel.flip(2).permute(1, 2, 0)
        k_taps = taps
    elif mode == 2:    # row t*N + c, k = j <- W[c, j, t]
        full = torch.zeros(rows_pad, 1, kp, device=dev)
        full[:taps * N, 0, :Cin] = w.permute(2, 0, 1).reshape(taps * N, Cin)
        k_taps = 1
    else:
        raise _lib.RfkError("split precision: the tap-split data-gradient weights belong to the training path, which this mode "
                            "does not cover")
    out = _split3(full.reshape(rows_pad, k_taps * kp), k_taps, kp)
    out.rfk_cin = cin_real
    return out, kp


def _pack_weight(weight, mode, perm, rows, rows_pad, kp, ktot, cin_real):
    if SPLIT:
        return _pack_weight_split(weight, mode, perm, rows, rows_pad, kp, ktot, cin_real)
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    N, Cin, kh, kw = w.shape
    out = torch.empty(rows_pad, ktot, device=w.device, dtype=torch.bfloat16)
    p32 = _perm32(perm)
    call("rfk_pack_weight", w.data_ptr(), N, Cin, kh * kw, mode, _p(p32), rows, kp, out.data_ptr(), rows_pad, ktot,
         _stream())
    out.rfk_cin = cin_real
    if w.data_ptr() == weight.data_ptr():   # packed straight from the parameter's storage: refreshable in place (derived.py)
        out.rfk_pack = [w.data_ptr(), out.data_ptr(), _p(p32), N, Cin, kh * kw, mode, rows, kp, rows_pad, ktot]
        out.rfk_perm = p32
    return out, kp


def pack_conv_weight_folded(weight, logs, bias, in_perm=None):
    """[N, Cin, kh, kw] f32 + the ActNorm (logs, bias [1,N,1,1]) that follows the conv -> bf16 [pad16(N), taps*cin_pad + 16]:
    rows scaled by exp(logs), 16 extra K columns with the shift bias*exp(logs) as (hi, lo) bf16 words
    (rfk_pack_weight_folded; operand of coupling_nn_fused)."""
    w = weight.detach()
    lg, bs = logs.detach(), bias.detach()
    for t in (w, lg, bs):
        if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
            raise _lib.RfkError("pack_conv_weight_folded: contiguous fp32 CUDA parameters expected")
    N, Cin, kh, kw = w.shape
    kp = cin_pad(Cin)
    ktot = kh * kw * kp + 16
    rows_pad = pad_to(N, 16)
    out = torch.empty(rows_pad, ktot, device=w.device, dtype=torch.bfloat16)
    p32 = _perm32(in_perm)
    call("rfk_pack_weight_folded", w.data_ptr(), N, Cin, kh * kw, 4, _p(p32), N, kp, lg.data_ptr(), bs.data_ptr(), out.data_ptr(),
         rows_pad, ktot, _stream())
    out.rfk_cin = Cin
    if w.data_ptr() == weight.data_ptr() and lg.data_ptr() == logs.data_ptr() and bs.data_ptr() == bias.data_ptr():
        # refreshable in place after an optimizer step (derived.py; mode 4 of rfk_pack_weights_batched)
        out.rfk_pack = [w.data_ptr(), out.data_ptr(), _p(p32), N, Cin, kh * kw, 4, N, kp, rows_pad, ktot, lg.data_ptr(),
                        bs.data_ptr()]
        out.rfk_perm = p32
    return out, kp


def pack_dgrad_weight_scaled(weight, logs):
    """pack_dgrad_weight with row r (the forward conv's input channel r) scaled by exp(logs[r]), logs [1,Cin,1,1] = the
    ActNorm that produced the conv's input: operand of conv_gemm_actbwd(scale=None)."""
    w, lg = weight.detach(), logs.detach()
    for t in (w, lg):
        if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
            raise _lib.RfkError("pack_dgrad_weight_scaled: contiguous fp32 CUDA parameters expected")
    N, Cin, kh, kw = w.shape
    assert lg.numel() == Cin
    kp = cin_pad(N)
    ktot = kh * kw * kp
    rows_pad = pad_to(Cin, 16)
    out = torch.empty(rows_pad, ktot, device=w.device, dtype=torch.bfloat16)
    call("rfk_pack_weight_folded", w.data_ptr(), N, Cin, kh * kw, 5, None, Cin, kp, lg.data_ptr(), None, out.data_ptr(), rows_pad,
         ktot, _stream())
    out.rfk_cin = N
    if w.data_ptr() == weight.data_ptr() and lg.data_ptr() == logs.data_ptr():
        out.rfk_pack = [w.data_ptr(), out.data_ptr(), 0, N, Cin, kh * kw, 5, Cin, kp, rows_pad, ktot, lg.data_ptr(), 0]
        out.rfk_perm = None
    return out, kp


def pack_dgrad_weight(weight, out_perm=None):
    """Weights of the data-gradient conv: dX = conv(dY, Wd) with Wd[ci, co, ky, kx] = W[co, ci, k-1-ky, k-1-kx]; rows
    (the dgrad's output channels = the forward conv's input channels) optionally permuted into staging-buffer order."""
    N, Cin, kh, kw = weight.shape
    rows = Cin if out_perm is None else out_perm.numel()
    kp = cin_pad(N)
    return _pack_weight(weight, 1, out_perm, rows, pad_to(rows, 16), kp, kh * kw * kp, N)


# --------------------------------------------------------------------------------------
# workspace pool: NHWC bf16 staging buffers shared by all modules of one shape (one stream, in order)
# --------------------------------------------------------------------------------------
_WS = {}


def workspace(tag, shape, device, dtype=torch.bfloat16):
    key = (tag, tuple(shape), str(device), dtype)
    t = _WS.get(key)
    if t is None:
        t = torch.zeros(shape, device=device, dtype=dtype)
        _WS[key] = t
    return t


_LIVE_GRAPHS = []     # weak references to captured graphs (graphs.py): their kernels hold raw pointers into the pools above


def register_graph(obj):
    import weakref
    _LIVE_GRAPHS.append(weakref.ref(obj))


def clear_workspaces():
    """Release the staging-buffer pool.  Refuses while a captured CUDA graph is alive: its kernels were recorded with raw
    pointers into these buffers and a replay after the release would read freed memory."""
    _LIVE_GRAPHS[:] = [r for r in _LIVE_GRAPHS if r() is not None]
    if _LIVE_GRAPHS:
        raise _lib.RfkError(f"clear_workspaces(): {len(_LIVE_GRAPHS)} captured graph(s) still reference the workspaces; "
                            "delete them first")
    _WS.clear()


# --------------------------------------------------------------------------------------
# weight repacking (one kernel launch per weight, cached by the callers until the parameter changes)
# --------------------------------------------------------------------------------------
def pack_tap_split_weight(weight):
    """[C, Cin, 3, 3] -> the 1x1 weight [pad16(9C), cin_pad] whose row t*C + c is W[c, :, ky, kx], t = 3*ky+kx."""
    C, Cin, kh, kw = weight.shape
    kp = cin_pad(Cin)
    return _pack_weight(weight, 2, None, kh * kw * C, pad_to(kh * kw * C, 16), kp, kp, Cin)


def pack_conv_weight(weight, in_perm=None, row_perm=None, n_pad=None):
    """[N, Cin, kh, kw] f32 -> bf16 [n_pad, taps*cin_pad] with k = tap*cin_pad + c (tap = 3*ky + kx)."""
    if row_perm is None and n_pad is None and weight.is_cuda and weight.shape[1] > 0:
        N, Cin, kh, kw = weight.shape
        kp = cin_pad(Cin)
        return _pack_weight(weight, 0, in_perm, N, pad_to(N, 16), kp, kh * kw * kp, Cin)
    w = weight.detach().float()
    N, Cin, kh, kw = w.shape
    if in_perm is not None:
        w = w[:, in_perm]
    cin_pad_ = cin_pad(max(Cin, 1))
    w = w.permute(0, 2, 3, 1).reshape(N, kh * kw, Cin)
    if row_perm is not None:
        rows = row_perm.numel()
        full = torch.zeros(rows, kh * kw, Cin, device=w.device)
        valid = row_perm >= 0
        full[valid] = w[row_perm[valid]]
        w, N = full, rows
    n_pad = n_pad or pad_to(N, 16)
    if SPLIT:
        full = torch.zeros(n_pad, kh * kw, cin_pad_, device=w.device)
        full[:N, :, :Cin] = w
        out = _split3(full.reshape(n_pad, kh * kw * cin_pad_), kh * kw, cin_pad_)
        out.rfk_cin = Cin
        return out, cin_pad_
    out = torch.zeros(n_pad, kh * kw, cin_pad_, device=w.device, dtype=torch.bfloat16)
    out[:N, :, :Cin] = w.to(torch.bfloat16)
    out = out.reshape(n_pad, kh * kw * cin_pad_).contiguous()
    out.rfk_cin = Cin  # real input channels, for FLOP accounting
    return out, cin_pad_
