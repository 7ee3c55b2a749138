"""Functional fp32 CPU restatement of the reference's Glow flow path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference lines it restates (paths relative to the reference repository root).
All tensors are torch CPU float32 in NCHW; parameters come in as a mapping with
the reference's own ``state_dict`` keys plus a key prefix, so golden fixtures
made from the reference load without renaming.

The functions are deliberately stateless: the data-dependent ActNorm
initialisation is an explicit function (``actnorm_init``) instead of hidden
module state, and random draws (dequantisation noise, Split2d / prior samples)
are passed in as explicit ``eps`` tensors.
"""
import math

import torch
import torch.nn.functional as F

LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------
# indexing helpers (bit exact)
# ----------------------------------------------------------------------------
def squeeze2d(x, undo_squeeze=False):
    """Flow/glow_modules.py:298-310.  out[b,4c+2dy+dx,i,j] = in[b,c,2i+dy,2j+dx]."""
    B, C, H, W = x.shape
    if not undo_squeeze:
        out = x.new_empty(B, 4 * C, H // 2, W // 2)
        for dy in range(2):
            for dx in range(2):
                out[:, (2 * dy + dx)::4] = x[:, :, dy::2, dx::2]
        return out
    out = x.new_empty(B, C // 4, 2 * H, 2 * W)
    for dy in range(2):
        for dx in range(2):
            out[:, :, dy::2, dx::2] = x[:, (2 * dy + dx)::4]
    return out


def split_feature(t, type="split"):
    """Utils/utils.py:86-91.  'split' = channel halves, 'cross' = even / odd channels."""
    C = t.shape[1]
    if type == "split":
        return t[:, : C // 2], t[:, C // 2:]
    if type == "cross":
        return t[:, 0::2], t[:, 1::2]
    raise ValueError(type)


def batch_reduce(x):
    """Utils/utils.py:25-28.  Sum everything but the batch dimension."""
    return x.reshape(x.shape[0], -1).sum(-1)


def act_fun(x, non_lin):
    """Utils/modules.py:8-19.  ReLU or LeakyReLU(0.2)."""
    if non_lin == "relu":
        return torch.clamp_min(x, 0.0)
    if non_lin == "leakyrelu":
        return torch.where(x >= 0, x, 0.2 * x)
    raise AssertionError("Please specify a activation type from the set {relu,leakyrelu}")


# ----------------------------------------------------------------------------
# ActNorm  (Flow/glow_modules.py:10-54)
# ----------------------------------------------------------------------------
def actnorm_init(x):
    """Flow/glow_modules.py:26-31.  Returns (bias, logs), each [1,C,1,1].

    bias = -mean over (N,H,W); logs = log(1 / (std_unbiased + 1e-6)).
    """
    C = x.shape[1]
    flat = x.transpose(0, 1).reshape(C, -1).double()
    n = flat.shape[1]
    mean = flat.mean(1)
    var = ((flat - mean[:, None]) ** 2).sum(1) / (n - 1)
    std = var.sqrt()
    logs = torch.log(1.0 / (std.float() + 1e-6))
    return (-mean.float()).view(1, C, 1, 1), logs.view(1, C, 1, 1)


def actnorm(x, bias, logs, logdet=None, reverse=False):
    """Flow/glow_modules.py:38-54.  fwd (x+b)*exp(logs); rev x*exp(-logs)-b."""
    hw = x.shape[2] * x.shape[3]
    bias = bias.view(1, -1, 1, 1)
    logs = logs.view(1, -1, 1, 1)
    if not reverse:
        y = (x + bias) * torch.exp(logs)
        d = logs.sum() * hw
    else:
        y = x * torch.exp(-logs) - bias
        d = -logs.sum() * hw
    if logdet is not None:
        logdet = logdet + d
    return y, logdet


def batchnorm_flow(x, log_gamma, beta, running_mean, running_var, logdet=None, reverse=False, training=False,
                   momentum=0.1, eps=1e-5):
    """Flow/glow_modules.py:73-104.  Per-position (C,H,W) statistics over the batch dimension only.

    Training & forward: batch mean / biased variance (+eps); the running buffers become
    running*momentum + batch*(1-momentum) (the reference's inverted momentum convention) and are returned.
    Otherwise the running buffers are used.  dlogdet = sum(log_gamma - 0.5*log(var)) over (C,H,W).
    Returns (out, logdet, running_mean, running_var)."""
    if training and not reverse:
        mean = x.mean(0)
        var = (x - mean).pow(2).mean(0) + eps
        running_mean = running_mean * momentum + mean * (1 - momentum)
        running_var = running_var * momentum + var * (1 - momentum)
    else:
        mean, var = running_mean, running_var
    d = torch.sum(log_gamma - 0.5 * torch.log(var))
    if not reverse:
        out = torch.exp(log_gamma) * ((x - mean) / var.sqrt()) + beta
        if logdet is not None:
            logdet = logdet + d
    else:
        out = ((x - beta) / torch.exp(log_gamma)) * var.sqrt() + mean
        if logdet is not None:
            logdet = logdet - d
    return out, logdet, running_mean, running_var


# ----------------------------------------------------------------------------
# InvConv  (Flow/glow_modules.py:150-221)
# ----------------------------------------------------------------------------
def invconv_weight(sd, prefix, reverse=False):
    """Flow/glow_modules.py:178-207.  Returns (W [C,C], sum(log|det|) per pixel).

    LU form when ``prefix+'lower'`` is present: W = P (L*mask + I)(U*mask^T + diag(sign_s e^{log_s})).
    """
    if prefix + "weight" in sd:
        w = sd[prefix + "weight"].float()
        per_pixel = torch.linalg.slogdet(w)[1]
        return (torch.linalg.inv(w) if reverse else w), per_pixel
    lower = sd[prefix + "lower"].float()
    upper = sd[prefix + "upper"].float()
    log_s = sd[prefix + "log_s"].float()
    sign_s = sd[prefix + "sign_s"].float()
    p = sd[prefix + "p"].float()
    C = lower.shape[0]
    strict_lower = torch.tril(torch.ones(C, C), -1)
    lo = lower * strict_lower + torch.eye(C)
    up = upper * strict_lower.t() + torch.diag(sign_s * torch.exp(log_s))
    per_pixel = log_s.sum()
    if reverse:
        w = torch.linalg.inv(up) @ (torch.linalg.inv(lo) @ torch.linalg.inv(p))
    else:
        w = p @ (lo @ up)
    return w, per_pixel


def invconv(x, sd, prefix, logdet=None, reverse=False):
    """Flow/glow_modules.py:209-221.  z[b,o,p] = sum_i W[o,i] x[b,i,p]."""
    w, per_pixel = invconv_weight(sd, prefix, reverse)
    hw = x.shape[2] * x.shape[3]
    z = torch.einsum("oi,bihw->bohw", w, x)
    if logdet is not None:
        logdet = logdet + per_pixel * hw if not reverse else logdet - per_pixel * hw
    return z, logdet


# ----------------------------------------------------------------------------
# coupling network pieces  (Flow/glow_modules.py:106-147)
# ----------------------------------------------------------------------------
def _same_pad(w):
    return ((w.shape[2] - 1) // 2, (w.shape[3] - 1) // 2)


def conv2d_norm(x, sd, prefix, training=False):
    """Flow/glow_modules.py:123-147.  norm='actnorm': bias-free conv, then ActNorm fwd;
    norm='batchnorm' (keys norm_type.running_mean present): conv with bias, then nn.BatchNorm2d (eval mode: running
    buffers; ``training``: batch statistics, the buffers are left alone -- they do not enter the output)."""
    w = sd[prefix + "conv.weight"].float()
    if prefix + "norm_type.running_mean" in sd:
        y = F.conv2d(x, w, sd[prefix + "conv.bias"].float(), 1, _same_pad(w))
        if training:
            return F.batch_norm(y, None, None, sd[prefix + "norm_type.weight"].float(), sd[prefix + "norm_type.bias"].float(),
                                True, 0.1, 1e-5)
        return F.batch_norm(y, sd[prefix + "norm_type.running_mean"].float(), sd[prefix + "norm_type.running_var"].float(),
                            sd[prefix + "norm_type.weight"].float(), sd[prefix + "norm_type.bias"].float(), False, 0.1, 1e-5)
    y = F.conv2d(x, w, None, 1, _same_pad(w))
    y, _ = actnorm(y, sd[prefix + "norm_type.bias"].float(), sd[prefix + "norm_type.logs"].float())
    return y


def conv2d_zeros(x, sd, prefix):
    """Flow/glow_modules.py:106-121.  (conv(x)+bias) * exp(3*logs)."""
    w = sd[prefix + "conv.weight"].float()
    y = F.conv2d(x, w, sd[prefix + "conv.bias"].float(), 1, _same_pad(w))
    return y * torch.exp(sd[prefix + "logs"].float().view(1, -1, 1, 1) * 3.0)


def coupling_nn(h, sd, prefix, non_lin="relu"):
    """Flow/glow_modules.py:232-238.  conv3x3-ActNorm-act-conv1x1-ActNorm-act-Conv2dZeros."""
    h = act_fun(conv2d_norm(h, sd, prefix + "0."), non_lin)
    h = act_fun(conv2d_norm(h, sd, prefix + "2."), non_lin)
    return conv2d_zeros(h, sd, prefix + "4.")


def clamp_log_scale(s, clamp_type, scale=None, scale_shift=None):
    """Flow/glow_modules.py:252-268."""
    if clamp_type == "glow":
        return torch.log(torch.sigmoid(s + 2.0))
    if clamp_type == "softclamp":
        return 2.5 * 0.636 * torch.atan(s / 2.5)
    if clamp_type == "realnvp":
        return scale.view(1, -1, 1, 1) * torch.tanh(s) + scale_shift.view(1, -1, 1, 1)
    return s


def affine_coupling(x, condition, sd, prefix, logdet=None, reverse=False,
                    clamp_type="realnvp", non_lin="relu"):
    """Flow/glow_modules.py:270-291."""
    assert condition.shape[2:4] == x.shape[2:4], "condition and x in affine needs to match"
    z1, z2 = split_feature(x, "split")
    out = coupling_nn(torch.cat([z1, condition], 1), sd, prefix + "net.", non_lin)
    shift, raw = split_feature(out, "cross")
    ls = clamp_log_scale(raw, clamp_type, sd.get(prefix + "scale"), sd.get(prefix + "scale_shift"))
    if not reverse:
        z2 = (z2 + shift) * torch.exp(ls)
        if logdet is not None:
            logdet = logdet + ls.sum(dim=(1, 2, 3))
    else:
        z2 = z2 * torch.exp(-ls) - shift
        if logdet is not None:
            logdet = logdet - ls.sum(dim=(1, 2, 3))
    return torch.cat([z1, z2], 1), logdet


def _normal_log_prob(z, mean, std):
    return -((z - mean) ** 2) / (2.0 * std * std) - torch.log(std) - 0.5 * LOG_2PI


def split2d(x, condition, sd, prefix, logdet=None, reverse=False, temperature=None,
            make_conditional=True, clamp_function="softplus", eps=None):
    """Flow/glow_modules.py:346-369.  ``eps`` replaces the reverse direction's N(0,1) draw."""
    if not reverse:
        z1, z2 = split_feature(x, "split")
    else:
        z1 = x
    if make_conditional:
        c = act_fun(conv2d_norm(condition, sd, prefix + "convcond.0."), "relu")
        c = act_fun(conv2d_norm(c, sd, prefix + "convcond.2."), "relu")
        h = torch.cat([z1, c], 1)
    else:
        h = z1
    mean, raw = split_feature(conv2d_zeros(h, sd, prefix + "conv.0."), "cross")
    if clamp_function == "softplus":
        std = F.softplus(raw) + 1e-8
    elif clamp_function == "exp":
        std = torch.exp(raw)
    else:
        raise AssertionError("Please specify a clamp function for the split2d from the set {softplus, exp}")
    if not reverse:
        if logdet is not None:
            logdet = logdet + _normal_log_prob(z2, mean, std).sum(dim=(1, 2, 3))
        return z1, logdet
    z2 = mean + std * temperature * eps
    return torch.cat([z1, z2], 1), logdet


# ----------------------------------------------------------------------------
# GlowStep / ListGlow  (Flow/glow.py)
# ----------------------------------------------------------------------------
def glow_step(x, condition, sd, prefix, logdet=None, reverse=False,
              clamp_type="realnvp", non_lin="relu", training=False):
    """Flow/glow.py:31-41; flow_norm='batchnorm' (keys norm.log_gamma present) in eval mode uses the running buffers,
    with ``training`` the batch statistics (Flow/glow_modules.py:73-87)."""
    if prefix + "norm.log_gamma" in sd:
        bn = lambda t, ld, rev: batchnorm_flow(t, sd[prefix + "norm.log_gamma"].float(), sd[prefix + "norm.beta"].float(),  # noqa: E731
                                               sd[prefix + "norm.running_mean"].float(),
                                               sd[prefix + "norm.running_var"].float(), ld, rev, training)[:2]
        if not reverse:
            x, logdet = bn(x, logdet, False)
            x, logdet = invconv(x, sd, prefix + "invconv.", logdet, False)
            return affine_coupling(x, condition, sd, prefix + "affine.", logdet, False, clamp_type, non_lin)
        x, logdet = affine_coupling(x, condition, sd, prefix + "affine.", logdet, True, clamp_type, non_lin)
        x, logdet = invconv(x, sd, prefix + "invconv.", logdet, True)
        return bn(x, logdet, True)
    if not reverse:
        x, logdet = actnorm(x, sd[prefix + "norm.bias"].float(), sd[prefix + "norm.logs"].float(), logdet, False)
        x, logdet = invconv(x, sd, prefix + "invconv.", logdet, False)
        x, logdet = affine_coupling(x, condition, sd, prefix + "affine.", logdet, False, clamp_type, non_lin)
    else:
        x, logdet = affine_coupling(x, condition, sd, prefix + "affine.", logdet, True, clamp_type, non_lin)
        x, logdet = invconv(x, sd, prefix + "invconv.", logdet, True)
        x, logdet = actnorm(x, sd[prefix + "norm.bias"].float(), sd[prefix + "norm.logs"].float(), logdet, True)
    return x, logdet


def listglow_layout(L, K):
    """Flow/glow.py:61-76: module order of ``glow_frame`` as (kind, level) tuples."""
    order = []
    for l in range(L):
        order.append(("squeeze", l))
        order.extend(("step", l) for _ in range(K))
        if l < L - 1:
            order.append(("split", l))
    return order


def listglow_f(x, condition, sd, L, K, logdet=0.0, prefix="", clamp_type="realnvp",
               non_lin="relu", make_conditional=True, split2d_act="softplus", training=False):
    """Flow/glow.py:105-117.  x -> z."""
    z = x
    for i, (kind, l) in enumerate(listglow_layout(L, K)):
        p = f"{prefix}glow_frame.{i}."
        if kind == "squeeze":
            z = squeeze2d(z, False)
        elif kind == "split":
            z, logdet = split2d(z, condition[l], sd, p, logdet, False, None, make_conditional, split2d_act)
        else:
            z, logdet = glow_step(z, condition[l], sd, p, logdet, False, clamp_type, non_lin, training)
    return z, logdet


def listglow_g(z, condition, sd, L, K, logdet=None, temperature=1.0, prefix="",
               clamp_type="realnvp", non_lin="relu", make_conditional=True,
               split2d_act="softplus", eps_list=None):
    """Flow/glow.py:90-102.  z -> x.  ``eps_list[l]`` is the N(0,1) draw of level l's Split2d."""
    x = z
    layout = listglow_layout(L, K)
    for i in reversed(range(len(layout))):
        kind, l = layout[i]
        p = f"{prefix}glow_frame.{i}."
        if kind == "squeeze":
            x = squeeze2d(x, True)
        elif kind == "split":
            x, logdet = split2d(x, condition[l], sd, p, logdet, True, temperature,
                                make_conditional, split2d_act, eps_list[l])
        else:
            x, logdet = glow_step(x, condition[l], sd, p, logdet, True, clamp_type, non_lin)
    return x, logdet


def listglow_prior(base_condition, sd, n, z_shape, learn_prior=True, non_lin="relu", prefix="", training=False):
    """Flow/glow.py:133-137.  Returns (mean, log_scale) as 'split' halves of the prior net."""
    if learn_prior:
        h = act_fun(conv2d_norm(base_condition, sd, prefix + "prior.0.", training), non_lin)
        h = act_fun(conv2d_norm(h, sd, prefix + "prior.2.", training), non_lin)
        out = conv2d_zeros(h, sd, prefix + "prior.4.")
    else:
        out = torch.zeros(n, 2 * z_shape[0], z_shape[1], z_shape[2])
    return split_feature(out, "split")


def listglow_log_prob(x, condition, base_condition, sd, L, K, n_bits, noise=None, logdet=0.0,
                      learn_prior=True, prefix="", **kw):
    """Flow/glow.py:119-141.  ``noise`` is the U(0,1/2^n_bits) dequantisation draw (or None)."""
    b, c, h, w = x.shape
    if noise is not None:
        x = x + noise
    obj_unif = -math.log(2.0 ** n_bits) * (c * h * w) * torch.ones(b)
    z, obj = listglow_f(x, condition, sd, L, K, logdet, prefix, **kw)
    obj = obj + obj_unif
    mean, log_scale = listglow_prior(base_condition, sd, b, z.shape[1:], learn_prior,
                                     kw.get("non_lin", "relu"), prefix, kw.get("training", False))
    obj = obj + batch_reduce(_normal_log_prob(z, mean, torch.exp(log_scale)))
    return z, -obj


def listglow_sample(condition, base_condition, sd, L, K, eps_prior, eps_list, temperature=0.8,
                    learn_prior=True, prefix="", **kw):
    """Flow/glow.py:143-160 with the prior draw and the Split2d draws made explicit."""
    n = eps_prior.shape[0]
    mean, log_scale = listglow_prior(base_condition, sd, n, eps_prior.shape[1:], learn_prior,
                                     kw.get("non_lin", "relu"), prefix)
    z = mean + torch.exp(log_scale) * temperature * eps_prior
    x, _ = listglow_g(z, condition, sd, L, K, None, temperature, prefix, eps_list=eps_list, **kw)
    return x


def bits_per_dim(nll, chw):
    """RFN/trainer.py:211-214 restricted to the flow term: nll / (ln2 * C*H*W)."""
    return nll / (math.log(2.0) * chw)
