"""Glow primitives with the reference's module interface, running on librfk's sm_100a kernels.

Mirrors ``Flow/glow_modules.py`` of cdglissov/recurrent-flows-msc: same class names,
constructor arguments, ``forward`` signatures and ``state_dict`` keys, so a
checkpoint of the reference loads unchanged.  The arithmetic is not PyTorch's: every forward
enqueues hand-written CUDA kernels through the C ABI (include/rfk.h).  There is no CPU path.

Training: ``ListGlow.log_prob`` (and ``ConvLSTM``) run under autograd through a tape-recording forward and a
hand-written backward (Flow/training.py, Utils/training.py).  The individual modules below, called on their own
with autograd recording enabled, refuse to run instead of silently returning tensors without a graph.
"""
import os

import torch
import torch.nn as nn

from .. import derived, ops
from ..Utils.utils import split_feature  # noqa: F401  (re-exported like the reference)
from ..Utils.modules import ActFun


FUSE_CONV2_TAPS = True   # AffineCoupling: fuse net.2 (1x1 conv + ActNorm + act) with the tap-split net.4 when 9*C <= 128
FUSE_COUPLING_NN = os.environ.get("RFK_FUSE_NN", "1") != "0"   # AffineCoupling: all three convs in ONE kernel (h1 and h2 in tensor memory)
FUSE_NN_MIN_TILES = int(os.environ.get("RFK_FUSE_NN_MIN_TILES", "48"))  # ... when there are at least this many 128-pixel tiles (below that the per-layer split-K launches win)
# ... and at most this many tap planes 9*C: the kernel takes up to 256 (two passes of its last GEMM), but beyond 128 it only
# matched the per-layer launches (config J level 3: forward 5.77 vs 5.70 ms; config D level 2: no change)
FUSE_NN_MAX_PLANES = int(os.environ.get("RFK_FUSE_NN_MAX_PLANES", "128"))
TAP_SPLIT_MAX_N = 2304  # AffineCoupling: tap-split form of the last conv up to C = 256 (K drops from 9*256 to 256)


def _require_no_grad():
    if torch.is_grad_enabled():
        raise RuntimeError("recurrent-flows-msc_b200: this module has no stand-alone autograd path; call it under "
                           "torch.no_grad() (density evaluation / sampling) or train through ListGlow.log_prob / ConvLSTM, "
                           "which run the hand-written backward")


_PARAM_EPOCH = [0]


def invalidate_caches():
    """Drop every tensor derived from parameters (packed weights, folded ActNorm/InvConv matrices).  Needed only when
    parameters were written through raw pointers (the fused Adam kernel, a CUDA-graph replay of a training step):
    in-place torch ops, load_state_dict and .to() are detected through the tensors' version counters."""
    _PARAM_EPOCH[0] += 1


def _ver_of(params):
    return (_PARAM_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in params)


class _Versioned:
    """Cache of tensors derived from parameters, rebuilt when a parameter is modified in place
    (optimizer step, load_state_dict) or re-allocated (.to(), .cuda()), or after invalidate_caches().

    ``refresh`` (optional callback(cache, key, params, slot)) registers a freshly built entry with derived.REFRESHER, which
    rewrites such entries IN PLACE with one batched kernel per kind after an optimizer step and re-stamps them, so that a
    training loop never takes the rebuild path again."""

    def __init__(self):
        self._store = {}

    def get(self, key, params, build, refresh=None):
        ver = _ver_of(params)
        hit = self._store.get(key)
        if hit is None or hit[0] != ver:
            with torch.no_grad():
                hit = [ver, build()]
            self._store[key] = hit
            if refresh is not None:
                refresh(self, key, params, hit)
        return hit[1]

    def clear(self):
        self._store.clear()


def _ld_begin(logdet, B, device, inplace=False):
    """Normalise the reference's logdet argument (None | number | 0-dim | [B]) to a [B] f32
    accumulator the kernels add into, plus whatever must be added back afterwards."""
    if logdet is None:
        return None, None
    if torch.is_tensor(logdet) and logdet.dim() == 1 and logdet.shape[0] == B and logdet.is_cuda:
        buf = logdet.detach()
        if buf.dtype != torch.float32 or not buf.is_contiguous() or not inplace:
            buf = buf.to(torch.float32).clone()
        return buf, None
    return torch.zeros(B, device=device, dtype=torch.float32), logdet


def _ld_end(buf, extra):
    if buf is None:
        return None
    if extra is None or (not torch.is_tensor(extra) and extra == 0):
        return buf
    return buf + extra


class _Ctx:
    """Per-level state ListGlow threads through its modules: the coupling network's NHWC input
    buffer with the condition already packed, and which other parts are already in place."""

    def __init__(self, nn_in, cond_channels):
        self.nn_in = nn_in
        self.cond_channels = cond_channels
        self.z1_packed = False
        self.pending = None   # (coupling module, taps, z, logdet buffer): a forward coupling whose tail is not applied yet

    def flush(self):
        """Apply a pending coupling tail in place (when no 1x1 mix follows that could absorb it); returns its tensor."""
        if self.pending is None:
            return None
        coupling, taps, z, ld = self.pending
        self.pending = None
        coupling.finish("taps", taps, z, ld, False)
        return z


# ----------------------------------------------------------------------------------------
class ActNorm(nn.Module):
    """Flow/glow_modules.py:10-54."""

    def __init__(self, num_channels):
        super().__init__()
        size = [1, num_channels, 1, 1]
        self.register_parameter("bias", nn.Parameter(torch.zeros(*size), requires_grad=True))
        self.register_parameter("logs", nn.Parameter(torch.zeros(*size), requires_grad=True))
        self.register_buffer("initialized", torch.tensor(0, dtype=torch.uint8))
        self._init_known = None  # host mirror of `initialized`: no .item() sync per call
        self._cache = _Versioned()

    def _load_from_state_dict(self, *a, **k):
        super()._load_from_state_dict(*a, **k)
        self._init_known = None
        self._cache.clear()

    def is_initialized(self):
        if self._init_known is None:
            self._init_known = bool(self.initialized.item() != 0)  # one sync, then cached
        return self._init_known

    def mark_initialized(self):
        self.initialized.fill_(1)
        self._init_known = True

    def initialize(self, input):
        """Flow/glow_modules.py:22-31: only in training mode; unbiased std, +1e-6."""
        if not self.training:
            return
        with torch.no_grad():
            ops.actnorm_init(ops.f32c(input), self.bias.data, self.logs.data)
            self.bias.add_(0.0)   # the kernel wrote through raw pointers: bump the version counters so that every
            self.logs.add_(0.0)   # cache derived from these parameters (here and in GlowStep) is rebuilt
        self._cache.clear()

    def maybe_initialize(self, input):
        if not self.is_initialized():
            self.initialize(input)
            self.mark_initialized()

    def affine(self):
        """(scale, shift) with y = x*scale + shift == (x + bias)*exp(logs); cached."""
        def build():
            s = torch.exp(self.logs.detach().float().reshape(-1))
            return s.contiguous(), (self.bias.detach().float().reshape(-1) * s).contiguous()
        return self._cache.get("affine", (self.bias, self.logs), build, derived.reg_affine(self.logs, self.bias, 1))

    def forward(self, input, logdet, reverse):
        _require_no_grad()
        self.maybe_initialize(input)
        x = ops.f32c(input)
        y = ops.actnorm(x, self.bias.data, self.logs.data, reverse)
        if logdet is not None:
            dlogdet = torch.sum(self.logs.detach()) * (x.shape[2] * x.shape[3])
            logdet = logdet - dlogdet if reverse else logdet + dlogdet
        return y, logdet


class Conv2dZeros(nn.Module):
    """Flow/glow_modules.py:106-121: zero-initialised conv, output * exp(3*logs)."""

    def __init__(self, in_channel, out_channel, kernel_size=[3, 3], stride=[1, 1]):
        super().__init__()
        assert list(stride) == [1, 1], "only stride 1 is used on this path"
        padding = (kernel_size[0] - 1) // 2
        self.conv = nn.Conv2d(in_channel, out_channel, kernel_size, stride, padding)
        self.logscale_factor = 3
        self.register_parameter("logs", nn.Parameter(torch.zeros(out_channel, 1, 1)))
        self.conv.weight.data.zero_()
        self.conv.bias.data.zero_()
        self.taps = kernel_size[0] * kernel_size[1]
        assert self.taps in (1, 9), "kernel must be 1x1 or 3x3"
        self._cache = _Versioned()

    def packed(self, key="id", in_perm=None):
        return self._cache.get(("w", key), (self.conv.weight,), lambda: ops.pack_conv_weight(self.conv.weight, in_perm), derived.reg_pack)

    def packed_dgrad(self, key="id", out_perm=None):
        """Weights of the data-gradient convolution (flipped taps, in/out channels swapped), rows in staging order."""
        return self._cache.get(("wd", key), (self.conv.weight,), lambda: ops.pack_dgrad_weight(self.conv.weight, out_perm), derived.reg_pack)

    def packed_dgrad_scaled(self, prev_logs):
        """Data-gradient weights with the scale of the ActNorm that produced this conv's input folded into the rows
        (ops.pack_dgrad_weight_scaled): conv_gemm_actbwd then needs no per-channel factor in its epilogue."""
        return self._cache.get(("wds", prev_logs.data_ptr()), (self.conv.weight, prev_logs),
                               lambda: ops.pack_dgrad_weight_scaled(self.conv.weight, prev_logs), derived.reg_pack)

    def packed_dgrad_taps(self, key="id", out_perm=None):
        """Tap-split form of the data-gradient weights (3x3 convs with few input channels: one GEMM with N = 9*Cin)."""
        return self._cache.get(("wd9", key), (self.conv.weight,), lambda: ops.pack_dgrad_taps_weight(self.conv.weight, out_perm), derived.reg_pack)

    def packed_taps(self):
        """Tap-split 1x1 form of the 3x3 weight (ops.pack_tap_split_weight), cached."""
        return self._cache.get(("w9",), (self.conv.weight,), lambda: ops.pack_tap_split_weight(self.conv.weight), derived.reg_pack)

    def affine(self):
        def build():
            s = torch.exp(self.logs.detach().float().reshape(-1) * self.logscale_factor)
            return s.contiguous(), (self.conv.bias.detach().float() * s).contiguous()
        return self._cache.get("affine", (self.logs, self.conv.bias), build,
                               derived.reg_affine(self.logs, self.conv.bias, self.logscale_factor))

    def fused(self, act, out, key="id", in_perm=None):
        """act NHWC bf16 -> out NCHW f32 = (conv + bias) * exp(3 logs)."""
        wgt, cin_pad = self.packed(key, in_perm)
        scale, shift = self.affine()
        return ops.conv_gemm(act, cin_pad, wgt, self.conv.out_channels, self.taps, scale, shift, "none", out)

    def forward(self, input):
        _require_no_grad()
        x = ops.f32c(input)
        B, C, H, W = x.shape
        act = ops.workspace(("cz_in", C), (B, H, W, ops.buf_ld(C)), x.device)
        ops.pack_nhwc(x, 0, C, act, 0)
        out = torch.empty(B, self.conv.out_channels, H, W, device=x.device, dtype=torch.float32)
        return self.fused(act, out)


class Conv2dNorm(nn.Module):
    """Flow/glow_modules.py:123-147: conv followed by ActNorm (bias-free conv) or, for norm='batchnorm', by
    nn.BatchNorm2d (conv with zero-initialised bias).  Either normalisation is a per-channel affine in the conv epilogue."""

    def __init__(self, in_channels, out_channels, kernel_size=[3, 3], stride=[1, 1], norm="actnorm"):
        super().__init__()
        assert list(stride) == [1, 1], "only stride 1 is used on this path"
        padding = [(kernel_size[0] - 1) // 2, (kernel_size[1] - 1) // 2]
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=(norm != "actnorm"))
        self.conv.weight.data.normal_(mean=0.0, std=0.05)
        self.norm = norm
        if self.norm == "actnorm":
            self.norm_type = ActNorm(out_channels)
        elif self.norm == "batchnorm":
            self.conv.bias.data.zero_()
            self.norm_type = nn.BatchNorm2d(out_channels)
        self.taps = kernel_size[0] * kernel_size[1]
        assert self.taps in (1, 9), "kernel must be 1x1 or 3x3"
        self._cache = _Versioned()

    def packed(self, key="id", in_perm=None):
        return self._cache.get(("w", key), (self.conv.weight,), lambda: ops.pack_conv_weight(self.conv.weight, in_perm), derived.reg_pack)

    def packed_dgrad(self, key="id", out_perm=None):
        """Weights of the data-gradient convolution (flipped taps, in/out channels swapped), rows in staging order."""
        return self._cache.get(("wd", key), (self.conv.weight,), lambda: ops.pack_dgrad_weight(self.conv.weight, out_perm), derived.reg_pack)

    def packed_dgrad_scaled(self, prev_logs):
        """Data-gradient weights with the scale of the ActNorm that produced this conv's input folded into the rows
        (ops.pack_dgrad_weight_scaled): conv_gemm_actbwd then needs no per-channel factor in its epilogue."""
        return self._cache.get(("wds", prev_logs.data_ptr()), (self.conv.weight, prev_logs),
                               lambda: ops.pack_dgrad_weight_scaled(self.conv.weight, prev_logs), derived.reg_pack)

    def packed_dgrad_taps(self, key="id", out_perm=None):
        """Tap-split form of the data-gradient weights (3x3 convs with few input channels: one GEMM with N = 9*Cin)."""
        return self._cache.get(("wd9", key), (self.conv.weight,), lambda: ops.pack_dgrad_taps_weight(self.conv.weight, out_perm), derived.reg_pack)

    def packed_folded(self, key="id", in_perm=None):
        """Forward weights with this layer's (initialised) ActNorm folded in: rows scaled by exp(logs), the shift in 16 extra K
        columns (ops.pack_conv_weight_folded) -- the operand form of the one-kernel coupling network."""
        an = self.norm_type
        return self._cache.get(("wf", key), (self.conv.weight, an.logs, an.bias),
                               lambda: ops.pack_conv_weight_folded(self.conv.weight, an.logs, an.bias, in_perm), derived.reg_pack)

    def foldable(self):
        """True when packed_folded applies: an ActNorm whose statistics are already known (no pending data-dependent init)."""
        return self.norm == "actnorm" and self.norm_type.is_initialized() and self.conv.weight.is_cuda

    def ready_for_fusion(self):
        """True when the per-channel affine is known without looking at the data (no pending ActNorm init, no
        training-mode batch statistics)."""
        if self.norm == "actnorm":
            return self.norm_type.is_initialized()
        return not (self.norm == "batchnorm" and self.norm_type.training)

    def _bn_eval_affine(self):
        bn = self.norm_type
        def build():
            s = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
            t = (self.conv.bias.detach().float() - bn.running_mean.float()) * s + bn.bias.detach().float()
            return s.contiguous(), t.contiguous()
        return self._cache.get("bn", (bn.weight, bn.bias, bn.running_mean, bn.running_var, self.conv.bias), build)

    def affine(self, act=None, wgt=None, cin_pad=None, keep=None):
        """(scale, shift) of the normalisation as a conv epilogue; data-dependent cases run a raw conv pass first.
        ``keep`` (dict): in training-mode batch norm, receives the raw conv output and the batch statistics the backward
        needs (Flow/training.py)."""
        n = self.conv.out_channels
        if self.norm == "actnorm":
            an = self.norm_type
            if not an.is_initialized():
                if an.training:  # data-dependent init on the raw convolution output (glow_modules.py:140-142)
                    B, H, W, _ = act.shape
                    raw = torch.empty(B, n, H, W, device=act.device, dtype=torch.float32)
                    ops.conv_gemm(act, cin_pad, wgt, n, self.taps, None, None, "none", raw)
                    an.initialize(raw)
                an.mark_initialized()
            return an.affine()
        if self.norm == "batchnorm":
            bn = self.norm_type
            if not bn.training:
                return self._bn_eval_affine()
            # training mode: batch statistics of (conv + bias), biased variance for the normalisation, unbiased for
            # the running estimate (torch.nn.BatchNorm2d semantics)
            B, H, W, _ = act.shape
            raw = torch.empty(B, n, H, W, device=act.device, dtype=torch.float32)
            ops.conv_gemm(act, cin_pad, wgt, n, self.taps, None, self.conv.bias.detach().float(), "none", raw)
            mean = torch.empty(n, device=act.device)
            std = torch.empty(n, device=act.device)
            ops.channel_stats(raw, mean, std)
            cnt = B * H * W
            var_b = std * std * ((cnt - 1) / cnt)
            with torch.no_grad():
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
                bn.running_mean.mul_(1 - m).add_(mean * m)
                bn.running_var.mul_(1 - m).add_(std * std * m)
                bn.num_batches_tracked += 1
            s = bn.weight.detach().float() / torch.sqrt(var_b + bn.eps)
            t = (self.conv.bias.detach().float() - mean) * s + bn.bias.detach().float()
            if keep is not None:
                keep.update(raw=raw, mean=mean, var=var_b)
            return s.contiguous(), t.contiguous()
        return None, (None if self.conv.bias is None else self.conv.bias.detach().float())

    def fused(self, act, out, act_fn="none", key="id", in_perm=None, out_off=0, keep=None):
        """act NHWC bf16 -> out (NHWC bf16 at channel out_off, or NCHW f32) = act_fn(norm(conv(act)))."""
        wgt, cin_pad = self.packed(key, in_perm)
        n = self.conv.out_channels
        scale, shift = self.affine(act, wgt, cin_pad, keep)
        if out.dtype == torch.bfloat16:
            B, H, W, _ = act.shape
            k_split = ops.choose_k_split(B * H * W, self.taps, cin_pad)
            if k_split > 1:   # few pixel tiles, long K: spread K over the SMs (reduction fused into the kernel)
                return ops.conv_gemm_splitk_fused(act, cin_pad, wgt, n, self.taps, k_split, scale, shift, act_fn, out,
                                                  out_off)
        return ops.conv_gemm(act, cin_pad, wgt, n, self.taps, scale, shift, act_fn, out, out_off)

    def forward(self, input):
        _require_no_grad()
        x = ops.f32c(input)
        B, C, H, W = x.shape
        act = ops.workspace(("cn_in", C), (B, H, W, ops.buf_ld(C)), x.device)
        ops.pack_nhwc(x, 0, C, act, 0)
        out = torch.empty(B, self.conv.out_channels, H, W, device=x.device, dtype=torch.float32)
        return self.fused(act, out)


class InvConv(nn.Module):
    """Flow/glow_modules.py:150-221: invertible 1x1 convolution, optionally LU-parameterised."""

    def __init__(self, num_channels, LU_decomposed):
        super().__init__()
        w_shape = [num_channels, num_channels]
        w_init = torch.linalg.qr(torch.randn(*w_shape))[0]
        if not LU_decomposed:
            self.weight = nn.Parameter(torch.Tensor(w_init))
        else:
            p, lower, upper = torch.linalg.lu(w_init)
            s = torch.diag(upper)
            self.register_buffer("p", p)
            self.register_buffer("sign_s", torch.sign(s))
            self.lower = nn.Parameter(lower)
            self.log_s = nn.Parameter(torch.log(torch.abs(s)))
            self.upper = nn.Parameter(torch.triu(upper, 1))
            self.l_mask = torch.tril(torch.ones(w_shape), -1)
            self.eye = torch.eye(*w_shape)
        self.w_shape = w_shape
        self.LU_decomposed = LU_decomposed
        self._cache = _Versioned()

    def _params(self):
        return (self.lower, self.upper, self.log_s) if self.LU_decomposed else (self.weight,)

    def _consts(self, dev):
        """(l_mask, eye) on the parameters' device.  The reference keeps them as plain CPU attributes and moves them on
        every call (Flow/glow_modules.py:172-173,190-191); here the device copies are made once (also: no host-to-device
        copy may happen inside a CUDA-graph capture)."""
        hit = self.__dict__.get("_consts_dev")
        if hit is None or hit[0].device != dev:
            hit = (self.l_mask.to(dev), self.eye.to(dev))
            self.__dict__["_consts_dev"] = hit
        return hit

    def matrices(self):
        """(W, W^-1, per-pixel log|det W|) as device tensors; cached until a parameter changes.
        Follows Flow/glow_modules.py:188-205 (three triangular inverses in the LU form)."""
        def build():
            if not self.LU_decomposed:
                w = self.weight.detach().float()
                return w.contiguous(), torch.linalg.inv(w).contiguous(), torch.linalg.slogdet(w)[1]
            l_mask, eye = self._consts(self.lower.device)
            lower = self.lower.detach() * l_mask + eye
            u = self.upper.detach() * l_mask.transpose(0, 1).contiguous()
            u = u + torch.diag(self.sign_s * torch.exp(self.log_s.detach()))
            w = torch.matmul(self.p, torch.matmul(lower, u))
            w_inv = torch.matmul(torch.linalg.inv(u), torch.matmul(torch.linalg.inv(lower), torch.linalg.inv(self.p)))
            return w.contiguous(), w_inv.contiguous(), torch.sum(self.log_s.detach())
        return self._cache.get("m", self._params(), build)

    def weight_fwd(self):
        """(W, per-pixel log|det W|) without the inverses -- all the forward direction (density, training) needs; no
        factorisation with a host-side status check, so it may run inside a CUDA-graph capture (LU form)."""
        def build():
            if not self.LU_decomposed:
                w = self.weight.detach().float()
                return w.contiguous(), torch.linalg.slogdet(w)[1]
            l_mask, eye = self._consts(self.lower.device)
            lower = self.lower.detach() * l_mask + eye
            u = self.upper.detach() * l_mask.transpose(0, 1).contiguous()
            u = u + torch.diag(self.sign_s * torch.exp(self.log_s.detach()))
            return torch.matmul(self.p, torch.matmul(lower, u)).contiguous(), torch.sum(self.log_s.detach())
        return self._cache.get("mf", self._params(), build)

    def get_weight(self, input, reverse):
        b, c, h, w = input.shape
        W, W_inv, per_pixel = self.matrices()
        weight = W_inv if reverse else W
        return weight.view(self.w_shape[0], self.w_shape[1], 1, 1), per_pixel * h * w

    def forward(self, input, logdet, reverse):
        _require_no_grad()
        x = ops.f32c(input)
        weight, dlogdet = self.get_weight(x, reverse)
        z = ops.mix1x1(x, weight.view(self.w_shape))
        if logdet is not None:
            logdet = logdet - dlogdet if reverse else logdet + dlogdet
        return z, logdet


class AffineCoupling(nn.Module):
    """Flow/glow_modules.py:223-291.  Three tensor-core convolutions; the ActNorm+ReLU of the hidden
    layers and the whole coupling tail (cross split, clamp, affine, per-sample log-det) are epilogues."""

    def __init__(self, x_size, condition_size, hidden_units=256, non_lin='relu', clamp_type="realnvp"):
        super().__init__()
        Bx, Cx, Hx, Wx = x_size
        B, C, H, W = condition_size
        channels = Cx // 2 + C
        self.net = nn.Sequential(
            Conv2dNorm(channels, hidden_units),
            ActFun(non_lin),
            Conv2dNorm(hidden_units, hidden_units, kernel_size=[1, 1]),
            ActFun(non_lin),
            Conv2dZeros(hidden_units, Cx),
        )
        self.non_lin = non_lin
        self.hidden_units = hidden_units
        self.clamp_type = clamp_type if clamp_type in ("glow", "softclamp", "realnvp") else "none"
        if clamp_type == "realnvp":
            self.scale = nn.Parameter(torch.zeros(Cx // 2, 1, 1), requires_grad=True)
            self.scale_shift = nn.Parameter(torch.zeros(Cx // 2, 1, 1), requires_grad=True)
        self._half, self._cond = Cx // 2, C

    def _perm(self, device):
        # NHWC staging order is [condition | z1]; the reference's weight expects cat[z1, condition]
        hit = self.__dict__.get("_perm_cache")
        if hit is None or hit.device != device:
            half, cc = self._half, self._cond
            hit = torch.cat([torch.arange(half, half + cc, device=device), torch.arange(0, half, device=device)])
            self.__dict__["_perm_cache"] = hit
        return hit

    def tail_params(self):
        """(scale, shift, clamp_type, clamp_scale, clamp_shift) of the coupling tail: Conv2dZeros' affine and the clamp."""
        scale, shift = self.net[4].affine()
        cs = self.scale.detach().reshape(-1) if self.clamp_type == "realnvp" else None
        csh = self.scale_shift.detach().reshape(-1) if self.clamp_type == "realnvp" else None
        return scale, shift, self.clamp_type, cs, csh

    def run_nn(self, z, condition, _ctx):
        """The coupling network on z1 = z[:, :C/2] and the condition.  Returns ("taps", taps [B,9C,H,W]) when the last
        conv runs in tap-split form (the caller gathers the planes, alone or fused with a 1x1 mix), else ("h2", h2)."""
        B, C, H, W = z.shape
        half, cc = C // 2, condition.shape[1]
        dev = z.device
        if _ctx is None:
            nn_in = ops.workspace(("cpl_in", half + cc), (B, H, W, ops.buf_ld(half + cc)), dev)
            ops.pack_nhwc(ops.f32c(condition), 0, cc, nn_in, 0)
            ops.pack_nhwc(z, 0, half, nn_in, cc)
        else:
            nn_in = _ctx.nn_in
            if not _ctx.z1_packed:
                ops.pack_nhwc(z, 0, half, nn_in, cc)
        hp = ops.buf_ld(self.hidden_units)
        first, last, mid = self.net[0], self.net[4], self.net[2]
        tap_split = last.taps == 9 and 9 * C <= TAP_SPLIT_MAX_N
        # the whole network in one kernel (csrc/coupling_nn.cu) on the big levels: neither hidden tensor touches HBM
        if (FUSE_COUPLING_NN and FUSE_CONV2_TAPS and not ops.SPLIT and tap_split and mid.taps == 1 and 9 * C <= FUSE_NN_MAX_PLANES
                and self.hidden_units % 64 == 0 and self.hidden_units <= 256 and first.foldable() and mid.foldable()
                and ops.gemm_m_tiles(B, H, W) >= FUSE_NN_MIN_TILES):
            w1f, cin_pad1 = first.packed_folded("cz", self._perm(dev))
            w2f, _ = mid.packed_folded()
            wgt9, _ = last.packed_taps()
            taps = ops.workspace(("cpl_taps", C), (B, 9 * C, H, W), dev, torch.float32)
            ops.coupling_nn_fused(nn_in, cin_pad1, first.taps, w1f, self.hidden_units, w2f, self.non_lin, wgt9, 9 * C, taps)
            return "taps", taps
        h1 = ops.workspace(("cpl_h1", self.hidden_units), (B, H, W, hp), dev)
        first.fused(nn_in, h1, self.non_lin, "cz", self._perm(dev))
        # conv1x1 -> ActNorm -> act -> tap-split conv3x3 in one kernel (h2 stays in tensor memory) when the shapes allow
        b2b = (FUSE_CONV2_TAPS and not ops.SPLIT and tap_split and mid.taps == 1 and ops.pad_to(9 * C, 16) <= 128
               and self.hidden_units % 64 == 0 and self.hidden_units <= 256 and mid.ready_for_fusion())
        h2 = None
        if not b2b:
            h2 = ops.workspace(("cpl_h2", self.hidden_units), (B, H, W, hp), dev)
            mid.fused(h1, h2, self.non_lin)
        if not tap_split:
            return "h2", h2
        # few output channels: one 1x1 GEMM with N = 9*C (activations read once, not once per tap); the nine shifted
        # planes are gathered afterwards together with the coupling tail
        wgt9, cin_pad = last.packed_taps()
        taps = ops.workspace(("cpl_taps", C), (B, 9 * C, H, W), dev, torch.float32)
        if b2b:
            w2, cin_pad2 = mid.packed()
            s2, t2 = mid.affine()
            ops.conv1x1_taps_fused(h1, cin_pad2, w2, self.hidden_units, s2, t2, self.non_lin, wgt9, 9 * C, taps)
        else:
            ops.conv_gemm(h2, cin_pad, wgt9, 9 * C, 1, None, None, "none", taps)
        return "taps", taps

    def finish(self, kind, t, out, ld, reverse):
        """Apply the coupling tail in place on out[:, C/2:] from run_nn's result."""
        scale, shift, clamp_type, cs, csh = self.tail_params()
        if kind == "taps":
            ops.coupling_tail_taps(t, out, scale, shift, clamp_type, cs, csh, ld, reverse)
        else:
            last = self.net[4]
            wgt, cin_pad = last.packed()
            ops.conv_gemm_coupling(t, cin_pad, wgt, out.shape[1], last.taps, scale, shift, out, clamp_type, cs, csh, ld,
                                   reverse)

    def forward(self, x, condition, logdet, reverse, _ctx=None):
        _require_no_grad()
        assert condition.shape[2:4] == x.shape[2:4], "condition and x in affine needs to match"
        z = ops.f32c(x)
        kind, t = self.run_nn(z, condition, _ctx)
        out = z.clone() if _ctx is None else z   # module contract: inputs are never modified; ListGlow owns its intermediates
        ld, extra = _ld_begin(logdet, z.shape[0], z.device, inplace=_ctx is not None)
        self.finish(kind, t, out, ld, reverse)
        return out, _ld_end(ld, extra)


class Squeeze2d(nn.Module):
    """Flow/glow_modules.py:294-310 (bit exact)."""

    def forward(self, x, undo_squeeze):
        _require_no_grad()
        return ops.squeeze2d(ops.f32c(x), bool(undo_squeeze))


class Split2d(nn.Module):
    """Flow/glow_modules.py:312-369."""

    def __init__(self, x_size, condition_size, make_conditional=True, clamp_function='softplus'):
        super().__init__()
        self.make_conditional = make_conditional
        Bx, Cx, Hx, Wx = x_size
        non_lin = 'relu'
        if make_conditional:
            B, C, H, W = condition_size
            channels = Cx // 2 + C
            self.convcond = nn.Sequential(
                Conv2dNorm(C, C),
                ActFun(non_lin),
                Conv2dNorm(C, C, kernel_size=[1, 1]),
                ActFun(non_lin),
            )
            self._cond = C
        else:
            channels = Cx // 2
            self._cond = 0
        self.conv = nn.Sequential(Conv2dZeros(channels, Cx),)
        self._half = Cx // 2
        if clamp_function not in ('softplus', 'exp'):
            assert False, 'Please specify a clamp function for the split2d from the set {softplus, exp}'
        self.clamp_function = clamp_function

    def _perm(self, device):
        """Staging order [convcond(condition) | z1] -> the conv weight's cat[z1, h] order (None when unconditional);
        second element: the z1 rows alone (for the data gradient w.r.t. z1)."""
        hit = self.__dict__.get("_perm_cache")
        if hit is None or hit[1].device != device:
            half, cc = self._half, self._cond
            perm = None
            if self.make_conditional:
                perm = torch.cat([torch.arange(half, half + cc, device=device), torch.arange(0, half, device=device)])
            hit = (perm, torch.arange(0, half, device=device))
            self.__dict__["_perm_cache"] = hit
        return hit

    def _params(self, z1_src, condition, _ctx):
        """(mean, raw log-scale) tensor [B, 2*half, H, W] from z1 = first `half` channels of z1_src."""
        B, _, H, W = z1_src.shape
        half, cc, dev = self._half, self._cond, z1_src.device
        sp_in = ops.workspace(("sp_in", half + cc), (B, H, W, ops.buf_ld(half + cc)), dev)
        perm = None
        if self.make_conditional:
            if _ctx is None:
                cbuf = ops.workspace(("sp_c", cc), (B, H, W, ops.buf_ld(cc)), dev)
                ops.pack_nhwc(ops.f32c(condition), 0, cc, cbuf, 0)
            else:
                cbuf = _ctx.nn_in   # condition already packed at channels [0, cc)
            t1 = ops.workspace(("sp_t1", cc), (B, H, W, ops.buf_ld(cc)), dev)
            self.convcond[0].fused(cbuf, t1, "relu")
            self.convcond[2].fused(t1, sp_in, "relu")
            perm = self._perm(dev)[0]
        ops.pack_nhwc(z1_src, 0, half, sp_in, cc)
        params = torch.empty(B, 2 * half, H, W, device=dev, dtype=torch.float32)
        return self.conv[0].fused(sp_in, params, "cz", perm)

    def forward(self, x, condition, logdet, reverse, temperature=None, _ctx=None, eps=None):
        _require_no_grad()
        x = ops.f32c(x)
        B, _, H, W = x.shape
        half = self._half
        params = self._params(x, condition, _ctx)
        if not reverse:
            z1 = torch.empty(B, half, H, W, device=x.device, dtype=torch.float32)
            ops.copy_channels(x, 0, z1, 0, half)
            ld, extra = _ld_begin(logdet, B, x.device, inplace=_ctx is not None)
            if ld is not None:
                ops.gauss_logp(x, half, params, half, ops.PAIR_CROSS, self.clamp_function, ld)
            return z1, _ld_end(ld, extra)
        z = torch.empty(B, 2 * half, H, W, device=x.device, dtype=torch.float32)
        ops.copy_channels(x, 0, z, 0, half)
        if eps is None:  # the N(0,1) draw stays in torch's generator (last row)
            eps = torch.randn(B, half, H, W, device=x.device, dtype=torch.float32)
        ops.gauss_sample(ops.f32c(eps), params, half, ops.PAIR_CROSS, self.clamp_function, temperature, z, half)
        return z, logdet


class BatchNormFlow(nn.Module):
    """Flow/glow_modules.py:56-104 (RealNVP-style batch norm as a flow layer; flow_norm='batchnorm').

    Parameters and running buffers are per POSITION, [1,C,H,W]; statistics are taken over the batch dimension only.
    Keeps the reference's conventions: running = running*momentum + batch*(1-momentum), variance + eps, batch
    statistics only in training mode and only in the forward direction."""

    def __init__(self, x_size, momentum=0.1, eps=1e-5):
        super().__init__()
        Bx, Cx, Hx, Wx = x_size
        size = [1, Cx, Hx, Wx]
        self.log_gamma = nn.Parameter(torch.zeros(size))
        self.beta = nn.Parameter(torch.zeros(size))
        self.momentum = momentum
        self.eps = eps
        self.register_buffer('running_mean', torch.zeros(size))
        self.register_buffer('running_var', torch.ones(size))
        self._cache = _Versioned()

    def _affine(self, mean, var, reverse):
        lg, beta = self.log_gamma.detach()[0], self.beta.detach()[0]
        d = torch.sum(lg - 0.5 * torch.log(var))
        if not reverse:
            a = torch.exp(lg) / torch.sqrt(var)
            c = beta - mean * a
        else:
            a = torch.sqrt(var) / torch.exp(lg)
            c = mean - beta * a
        return a.contiguous(), c.contiguous(), d

    def forward(self, input, logdet, reverse):
        _require_no_grad()
        x = ops.f32c(input)
        if self.training and reverse == False:  # noqa: E712  (the reference's condition)
            mean, var = ops.batch_stats_pos(x, self.eps)
            self.batch_mean, self.batch_var = mean, var
            with torch.no_grad():
                self.running_mean.mul_(self.momentum).add_(mean * (1 - self.momentum))
                self.running_var.mul_(self.momentum).add_(var * (1 - self.momentum))
            a, c, d = self._affine(mean, var, False)
        else:
            a, c, d = self._cache.get(("aff", bool(reverse)), (self.log_gamma, self.beta, self.running_mean, self.running_var),
                                      lambda: self._affine(self.running_mean[0], self.running_var[0], bool(reverse)))
        z = ops.affine_pos(x, a, c)
        if logdet is not None:
            logdet = logdet - d if reverse else logdet + d
        return z, logdet
