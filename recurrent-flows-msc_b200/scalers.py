"""The conv stacks either side of the hot path -- the reference's VGG_downscaler (feature extractor), VGG_upscaler
(condition pyramid) and SimpleParamNet (prior / encoder), Utils/modules.py:43-244 -- on librfk's tcgen05 convolution kernel
in the INFERENCE direction: ``RFN.predict`` runs them once per predicted frame.

``accelerate_scalers(model)`` walks a built reference model and gives every such module a fast ``forward`` that is taken
when the module is in eval mode and autograd is off; otherwise (training) the module's own PyTorch forward runs, so
parameters, ``state_dict`` keys and gradients are untouched.  In the fast path each ``Conv2d(3x3 or 1x1, stride 1) ->
NormLayer -> activation`` triple is ONE ``rfk_conv_gemm`` launch: eval-mode BatchNorm2d (or no norm) and the conv bias
fold into the epilogue's per-channel (scale, shift), ReLU / LeakyReLU(0.2) is the epilogue activation, and consecutive
triples hand their activations over as NHWC bf16 without going back to NCHW fp32.  Pooling, nearest upsampling, squeeze,
concatenation with skip tensors, tanh and the softplus of SimpleParamNet stay in torch (tiny elementwise work); stride-2
convolutions, transposed convolutions and InstanceNorm fall back to the module's own layers.
"""
import types

import torch
import torch.nn as nn

from . import ops
from .Flow.glow_modules import _Versioned


def _act_kind(m):
    """(epilogue activation, torch post-op) of an activation module of the reference."""
    net = getattr(m, "net", None)                      # Utils.modules.ActFun wraps nn.ReLU / nn.LeakyReLU(0.2)
    if isinstance(net, nn.ReLU) or isinstance(m, nn.ReLU):
        return "relu", None
    if isinstance(net, nn.LeakyReLU) or isinstance(m, nn.LeakyReLU):
        slope = (net if net is not None else m).negative_slope
        return ("leakyrelu", None) if abs(slope - 0.2) < 1e-12 else ("none", m)
    return "none", m                                     # nn.Tanh, tanh0_5, anything else: applied by torch afterwards


class _Conv:
    """One conv (+ folded norm + epilogue activation)."""

    def __init__(self, conv, norm, act):
        self.conv, self.norm, self.act_mod = conv, norm, act
        self.act, self.post = _act_kind(act) if act is not None else ("none", None)
        self.taps = conv.kernel_size[0] * conv.kernel_size[1]
        self._cache = _Versioned()

    @staticmethod
    def supported(conv, norm):
        if not isinstance(conv, nn.Conv2d) or conv.stride != (1, 1) or conv.dilation != (1, 1) or conv.groups != 1:
            return False
        if conv.kernel_size not in ((3, 3), (1, 1)) or conv.padding != ((conv.kernel_size[0] - 1) // 2,) * 2:
            return False
        inner = getattr(norm, "norm", norm)
        return norm is None or isinstance(inner, nn.BatchNorm2d) or type(inner).__name__ == "NoNorm"

    def params(self):
        bn = getattr(self.norm, "norm", self.norm)
        ps = [self.conv.weight] + ([self.conv.bias] if self.conv.bias is not None else [])
        if isinstance(bn, nn.BatchNorm2d):
            ps += [t for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var) if t is not None]
        return tuple(ps)

    def packed(self):
        def build():
            wgt, cin_pad = ops.pack_conv_weight(self.conv.weight)
            n = self.conv.out_channels
            dev = self.conv.weight.device
            bn = getattr(self.norm, "norm", self.norm)
            bias = self.conv.bias.detach().float() if self.conv.bias is not None else torch.zeros(n, device=dev)
            if isinstance(bn, nn.BatchNorm2d):
                s = torch.rsqrt(bn.running_var.float() + bn.eps)
                if bn.weight is not None:
                    s = s * bn.weight.detach().float()
                t = (bias - bn.running_mean.float()) * s
                if bn.bias is not None:
                    t = t + bn.bias.detach().float()
            else:
                s, t = torch.ones(n, device=dev), bias
            return wgt, cin_pad, s.contiguous(), t.contiguous()
        return self._cache.get("w", self.params(), build)

    def __call__(self, act_nhwc, out):
        wgt, cin_pad, s, t = self.packed()
        return ops.conv_gemm(act_nhwc, cin_pad, wgt, self.conv.out_channels, self.taps, s, t, self.act, out)


def _plan(seq):
    """nn.Sequential of the reference's layer triples -> list of _Conv | nn.Module (run by torch on NCHW fp32)."""
    mods = list(seq.children())
    plan, i = [], 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Conv2d):
            norm = mods[i + 1] if i + 1 < len(mods) and hasattr(mods[i + 1], "norm") else None
            j = i + (2 if norm is not None else 1)
            act = mods[j] if j < len(mods) and not isinstance(mods[j], (nn.Conv2d, nn.MaxPool2d, nn.Upsample)) and \
                not hasattr(mods[j], "norm") and type(mods[j]).__name__ != "Squeeze2dDecoder" else None
            if _Conv.supported(m, norm):
                plan.append(_Conv(m, norm, act))
                i = j + (1 if act is not None else 0)
                continue
        plan.append(m)
        i += 1
    return plan


def _run(plan, x):
    """x NCHW fp32 -> NCHW fp32 through the planned stack."""
    k = 0
    while k < len(plan):
        step = plan[k]
        if not isinstance(step, _Conv):
            x = step(x)
            k += 1
            continue
        # a run of consecutive convs: NHWC bf16 in between, NCHW fp32 out of the last one
        run = [step]
        while k + len(run) < len(plan) and isinstance(plan[k + len(run)], _Conv) and run[-1].post is None:
            run.append(plan[k + len(run)])
        x = ops.f32c(x)
        B, C, H, W = x.shape
        cur = ops.workspace(("sc_in", C, H, W), (B, H, W, ops.buf_ld(C)), x.device)
        ops.pack_nhwc(x, 0, C, cur, 0)
        for idx, cv in enumerate(run):
            n = cv.conv.out_channels
            if idx + 1 < len(run):
                nxt = ops.workspace(("sc_mid", idx % 2, n, H, W), (B, H, W, ops.buf_ld(n)), x.device)
                cv(cur, nxt)
                cur = nxt
            else:
                x = cv(cur, torch.empty(B, n, H, W, device=x.device, dtype=torch.float32))
                if cv.post is not None:
                    x = cv.post(x)
        k += len(run)
    return x


def _fast_ok(module):
    return not module.training and not torch.is_grad_enabled()


def _downscaler_forward(self, x, block_size=None):
    if not _fast_ok(self):
        return self._rfk_orig_forward(x, block_size)
    outputs = []
    for i in range(self.L):
        x = _run(self._rfk_plans[i], x)
        if self.skip_con:
            outputs.append(x)
        else:
            outputs = x
    return outputs


def _upscaler_forward(self, x, skip_list=None):
    if not _fast_ok(self):
        return self._rfk_orig_forward(x, skip_list)
    outputs = []
    skips = list(reversed(skip_list)) if self.skips else None      # the reference reverses the caller's list in place, twice
    for i in range(self.L):
        if i > 0:
            x = self.upscales_nets[i - 1](x)
        if self.skips:
            x = _run(self._rfk_plans[i], torch.cat((x, skips[i]), dim=1))
        else:
            x = _run(self._rfk_plans[i], x)
        outputs.append(x)
    outputs.reverse()
    return outputs


def _paramnet_forward(self, x):
    if not _fast_ok(self):
        return self._rfk_orig_forward(x)
    out = _run(self._rfk_plans[0], x)
    loc, log_scale = _run(self._rfk_plans[1], out).chunk(2, 1)
    return loc, self.softplus(log_scale)


def accelerate_scalers(model):
    """Give every VGG_downscaler / VGG_upscaler / SimpleParamNet inside `model` the fast eval-mode forward.  Returns the
    names of the patched sub-modules.  Idempotent."""
    patched = []
    for name, m in model.named_modules():
        kind = type(m).__name__
        if hasattr(m, "_rfk_orig_forward"):
            continue
        if kind == "VGG_downscaler":
            plans, fwd = [_plan(net) for net in m.l_nets], _downscaler_forward
        elif kind == "VGG_upscaler":
            plans, fwd = [_plan(net) for net in m.l_nets], _upscaler_forward
        elif kind == "SimpleParamNet":
            plans, fwd = [_plan(m.net), _plan(nn.Sequential(m.param_net))], _paramnet_forward
        else:
            continue
        m.__dict__["_rfk_plans"] = plans
        m.__dict__["_rfk_orig_forward"] = m.forward
        m.forward = types.MethodType(fwd, m)
        patched.append(name or kind)
    return patched
