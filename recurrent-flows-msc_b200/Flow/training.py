"""Training path of ``ListGlow.log_prob``: reverse-mode differentiation of the flow on librfk's backward kernels.

The reference trains the flow through PyTorch autograd over ~186 ATen ops per GlowStep (Flow/glow.py:128-141 called
from RFN/RFN.py's loss).  Here the forward pass records a tape of the tensors each module's backward needs and the
backward pass walks it in reverse, launching hand-written kernels:

  coupling tail           rfk_coupling_taps_bwd   (gradients of the affine update, clamp and Conv2dZeros scale)
  3x3 / 1x1 convolutions  rfk_conv_wgrad (weight gradient, tensor cores) + rfk_conv_gemm on flipped weights (data gradient)
  ActNorm + activation    rfk_act_affine_bwd      (per-channel reductions give d logs / d bias)
  ActNorm . InvConv mix   rfk_mix1x1_wgrad + rfk_mix1x1 with the transposed matrix
  Split2d / prior density rfk_gauss_logp_bwd

Only the map from module parameters to the folded per-step quantities (C x C mix matrix from the LU factors and the
ActNorm scale, its log-determinant) is differentiated by torch autograd, on C x C tensors.  Activation gradients
between convolutions are bf16 (as the activations are); parameter gradients accumulate in fp32.

The batch-norm options of the reference (flow_norm='batchnorm': per-position BatchNormFlow; base_norm='batchnorm':
nn.BatchNorm2d in the prior) are differentiated with torch ops around the same conv kernels (small tensors, non-default).
Not covered (raise NotImplementedError): couplings with more than 256 channels (no tap-split form), gradients w.r.t. a
tensor-valued ``logdet`` argument.
"""
import os

import torch

from .. import ops
from .glow_modules import BatchNormFlow, Split2d, Squeeze2d, TAP_SPLIT_MAX_N, FUSE_NN_MIN_TILES, FUSE_NN_MAX_PLANES


DGRAD_TAP_SPLIT_MAX_N = 512         # tap-split data gradient when 9*Cin <= this ...
DGRAD_TAP_SPLIT_MIN_PIXELS = 32768  # ... and the launch has enough pixels to be bandwidth- rather than latency-bound
# Weight gradients are off the critical path of the sweep (nothing downstream reads them): they go to a second stream and
# overlap the data-gradient chain; in a captured training step this becomes a parallel branch of the CUDA graph.  Measured
# on the 570-frame step (one B200): levels 3-5 only (<= 72 pixel tiles on 148 SMs) 28.7 ms, + level 3 28.2, + level 2 27.5,
# all levels 27.1 -- the two kernel families complement each other (the weight gradients are MMA-issue / HBM bound, the
# data gradients epilogue bound) and fill each other's tail waves.  RFK_WGRAD_SIDE_STREAM=0 disables,
# RFK_WGRAD_SIDE_MAX_PIXELS limits it to launches of at most that many pixels.
WGRAD_SIDE_MAX_PIXELS = int(os.environ.get("RFK_WGRAD_SIDE_MAX_PIXELS", str(1 << 40)))
WGRAD_SIDE_STREAM = os.environ.get("RFK_WGRAD_SIDE_STREAM", "1") != "0"
# Hidden layers of the coupling network: fold the ActNorm + activation backward into the epilogue of the data-gradient GEMM
# that produces its input (rfk_conv_gemm_actbwd) and take the ActNorm parameter gradients from the layer's own weight
# gradient (rfk_actnorm_param_bwd) instead of a separate 6 B/element pass over the activations.  RFK_FUSE_ACT_BWD=0 disables.
FUSE_ACT_BWD = os.environ.get("RFK_FUSE_ACT_BWD", "1") != "0"


# e^{logs} of the ActNorm below folded into the data-gradient weight rows (conv_gemm_actbwd with scale = None).  Measured neutral
# (level 1: 160 / 166 us against 154 / 159 us -- that epilogue is not bound by its per-channel loads), so it stays opt-in.
FOLD_ACTBWD_SCALE = os.environ.get("RFK_FOLD_ACTBWD_SCALE", "0") == "1"
# Tap-split data gradient of the coupling network's first conv gathered straight into dz / the condition's gradient.
FUSE_GATHER_ACC = os.environ.get("RFK_FUSE_GATHER_ACC", "1") != "0"
# Recompute mode for every flow that does not set flow.recompute itself (see _glowstep_fwd).
RECOMPUTE = os.environ.get("RFK_RECOMPUTE", "0") == "1"
# Training forward of the coupling network as ONE kernel with h1 / h2 as side outputs (RFK_FUSE_NN_TRAIN=0: three launches).
FUSE_NN_TRAIN = os.environ.get("RFK_FUSE_NN_TRAIN", os.environ.get("RFK_FUSE_NN", "1")) != "0"


_SIDE = {}


def _side_stream(device):
    s = _SIDE.get(device)
    if s is None:
        s = _SIDE[device] = torch.cuda.Stream(device=device)
    return s


class _State:
    """What the backward sweep carries: the gradient of the current activation, the per-sample weight on the
    log-determinant, and the accumulated parameter / condition gradients."""

    def __init__(self, dz, g_ld, n_levels):
        self.dz = dz
        self.g_ld = g_ld
        self.G = g_ld.sum().reshape(1)
        self.dcond = [None] * n_levels
        self.dbase = None
        self.grads = {}
        self.direct = set()
        self.folds = []
        self.prior_announced = False
        self.recomputed = 0       # GlowSteps whose activations were regenerated in this sweep (recompute mode)
        self.side = None          # second stream for small weight-gradient launches
        self.side_keep = []       # tensors the side stream still reads (kept alive until the join)

    def wgrad(self, x_act, cin, da, n, taps, out, perm, after=None):
        """Weight gradient launch; small ones go to the side stream.  ``after``: work that consumes the weight gradient,
        enqueued right behind it on the same stream."""
        B, H, W, _ = da.shape
        if not WGRAD_SIDE_STREAM or B * H * W > WGRAD_SIDE_MAX_PIXELS:
            ops.conv_wgrad(x_act, cin, da, n, taps, out=out, perm=perm)
            if after is not None:
                after()
            return
        main = torch.cuda.current_stream()
        if self.side is None:
            self.side = _side_stream(da.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            ops.conv_wgrad(x_act, cin, da, n, taps, out=out, perm=perm, ws_slot=1)
            if after is not None:
                after()
        self.side_keep.append((x_act, da, out))

    def side_run(self, pixels, fn, keep):
        """Run fn() (launches whose results are only needed after the next join) on the side stream."""
        if not WGRAD_SIDE_STREAM or pixels > WGRAD_SIDE_MAX_PIXELS:
            fn()
            return
        main = torch.cuda.current_stream()
        if self.side is None:
            self.side = _side_stream(self.dz.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            fn()
        self.side_keep.append(keep)

    def join(self):
        """The main stream waits for the side stream; the tensors it was reading may be released afterwards."""
        if self.side is not None and self.side_keep:
            ev = torch.cuda.Event()
            ev.record(self.side)
            torch.cuda.current_stream().wait_event(ev)
            self.side_keep.clear()

    def out(self, param):
        """fp32 buffer (flat, zero or holding earlier contributions) the kernels ACCUMULATE the gradient of `param` into.
        For parameters whose .grad storage is owned by FlatAdam (views of its flat gradient buffer, marked _rfk_direct)
        that IS the .grad -- exactly autograd's accumulate semantics, without a per-parameter add or the 750-way
        concatenation afterwards; otherwise a zeroed slice of this sweep's arena, returned to autograd at the end."""
        g = param.grad
        if g is not None and getattr(param, "_rfk_direct", False) and g.dtype == torch.float32 and g.is_contiguous():
            self.direct.add(id(param))
            return g.view(-1)
        hit = self.grads.get(id(param))
        if hit is None:
            hit = self.grads[id(param)] = ops._zeros(param.numel(), param.device)
        return hit

    def add(self, param, g):
        """Legacy path: g is a tensor this sweep owns; added into the parameter's accumulation buffer."""
        self.out(param).add_(g.reshape(-1))

    def cond_acc(self, l, B, n, H, W):
        """The (zero-initialised) gradient of level l's condition, for kernels that accumulate into it in place."""
        if self.dcond[l] is None:
            self.dcond[l] = torch.zeros(B, n, H, W, device=self.dz.device, dtype=torch.float32)
        return self.dcond[l]

    def add_cond(self, l, src, n):
        """dcond[l] += src[:, :n] (src fp32 NCHW with the condition's channels first)."""
        if self.dcond[l] is None:
            B, _, H, W = src.shape
            self.dcond[l] = torch.zeros(B, n, H, W, device=src.device, dtype=torch.float32)
        ops.add_channels(self.dcond[l], 0, src, 0, n)

    def fold_buf(self, flow, step, C):
        """Zeroed scratch of one GlowStep's fold backward: dWf [C,C] | dbf [C] | A [C,C].  One persistent flat buffer per
        flow (pointer-stable, so the backward table is built once and the sweep can be captured in a CUDA graph), cleared
        by a single memset at the start of every sweep."""
        hit = flow.__dict__.get("_fold_flat")
        if hit is None or hit[0].device != self.dz.device:
            offs, total = {}, 0
            for m in flow.glow_frame:
                if hasattr(m, "invconv"):
                    c = m.invconv.w_shape[0]
                    offs[id(m)] = (total, 2 * c * c + c)
                    total += (2 * c * c + c + 3) // 4 * 4
            hit = (torch.zeros(total, device=self.dz.device, dtype=torch.float32), offs)
            flow.__dict__["_fold_flat"] = hit
            self.fold_cleared = True
        if not getattr(self, "fold_cleared", False):
            hit[0].zero_()
            self.fold_cleared = True
        off, n = hit[1][id(step)]
        return hit[0][off:off + n]


def _nhwc(B, H, W, c, dev):
    """NHWC bf16 staging buffer for c channels.  Pad columns must be finite zeros (they meet zero weights in the GEMMs);
    a buffer without pad columns is fully written by its producer and needs no fill (a 300 MB memset at level 1)."""
    ld = ops.cin_pad(c)
    return (torch.empty if ld == c else torch.zeros)(B, H, W, ld, device=dev, dtype=torch.bfloat16)


def _act_grad(h, act_fn):
    if act_fn == "relu":
        return (h > 0).to(h.dtype)
    if act_fn == "leakyrelu":
        return torch.where(h > 0, torch.ones_like(h), torch.full_like(h, 0.2))
    return torch.ones_like(h)


def _batchnorm_act_bwd(st, mod, dh, h, act_fn, keep):
    """Backward of h = act(BatchNorm2d(conv + bias)) in training mode (Flow/glow_modules.py:134-137,144-145; the
    reference's base_norm='batchnorm' option, used only by the prior on the coarsest maps): per-channel batch-norm backward
    over (B,H,W) written with torch ops on the fp32 raw conv output kept by the forward -- a handful of launches on
    [B,n,2,2] tensors.  Returns da (bf16 NHWC) for the conv's weight / data gradients."""
    bn, n = mod.norm_type, mod.conv.out_channels
    B, H, W, ld = dh.shape
    a = keep["raw"]                                                     # conv + bias, fp32 NCHW
    mean, var = keep["mean"].view(1, n, 1, 1), keep["var"].view(1, n, 1, 1)
    inv_std = torch.rsqrt(var + bn.eps)
    xhat = (a - mean) * inv_std
    dv = dh[..., :n].float().permute(0, 3, 1, 2) * _act_grad(h[..., :n].float().permute(0, 3, 1, 2), act_fn)
    gamma = bn.weight.detach().float().view(1, n, 1, 1)
    st.add(bn.weight, (dv * xhat).sum((0, 2, 3)))
    st.add(bn.bias, dv.sum((0, 2, 3)))
    m1 = dv.mean((0, 2, 3), keepdim=True)
    m2 = (dv * xhat).mean((0, 2, 3), keepdim=True)
    da32 = (gamma * inv_std * (dv - m1 - xhat * m2)).contiguous()
    st.add(mod.conv.bias, da32.sum((0, 2, 3)))
    da = _nhwc(B, H, W, n, dh.device)
    ops.pack_nhwc(da32, 0, n, da, 0)
    return da


def _dgrad_actbwd(st, conv_mod, da_in, prev_mod, h, act_fn):
    """Data gradient of conv_mod fused with the backward of h = act(ActNorm(prev_mod.conv(...))), the tensor conv_mod read:
    returns (da, colsum) = gradient w.r.t. prev_mod's raw conv output (bf16 NHWC) and its per-channel sums."""
    B, H, W, _ = da_in.shape
    n = prev_mod.conv.out_channels
    an = prev_mod.norm_type
    if FOLD_ACTBWD_SCALE and conv_mod.conv.weight.is_cuda and an.logs.dtype == torch.float32:
        # e^{logs} of the ActNorm below rides in the weight rows: the epilogue (bound by its shared-memory loads) reads no
        # per-channel factor
        wd, cp = conv_mod.packed_dgrad_scaled(an.logs)
        scale = None
    else:
        wd, cp = conv_mod.packed_dgrad("id", None)
        scale, _ = an.affine()
    da = torch.empty(B, H, W, n, device=da_in.device, dtype=torch.bfloat16)
    colsum = ops._zeros(n, da_in.device)
    ops.conv_gemm_actbwd(da_in, cp, wd, n, conv_mod.taps, scale, act_fn, h, da, colsum)
    return da, colsum


def _wgrad_actnorm(st, mod, x_act, cin, da, colsum, perm=None):
    """Weight gradient of a Conv2dNorm(ActNorm) layer into a zeroed scratch, then its ActNorm gradients from that scratch
    and the column sums of da (rfk_actnorm_param_bwd also adds the scratch into the weight's gradient buffer)."""
    w = mod.conv.weight
    n = mod.conv.out_channels
    k = 3 if mod.taps == 9 else 1
    dWp = ops._zeros(w.numel(), da.device)
    g_w, g_logs, g_bias = st.out(w), st.out(mod.norm_type.logs), st.out(mod.norm_type.bias)
    bias = mod.norm_type.bias.detach().reshape(-1)
    st.wgrad(x_act, cin, da, n, mod.taps, dWp.view(n, w.shape[1], k, k), perm,
             after=lambda: ops.actnorm_param_bwd(w.detach(), dWp, colsum, bias, g_w, g_logs, g_bias))


def _conv_bwd(st, mod, x_act, cin, da, perm=None, dgrad_out=None, key="id", skip_wgrad=False, acc=None):
    """Backward of mod.conv given da (bf16 NHWC gradient of the raw convolution output): weight gradient into the
    state, data gradient into dgrad_out (bf16 NHWC or fp32 NCHW; channels in the staging order of x_act).
    ``acc`` = (acc0, n0, acc1) instead of dgrad_out: the data gradient is ADDED in place, channels [0, n0) into acc0 and the
    rest into the first channels of acc1 (tap-split path only; returns False when that path does not apply)."""
    if acc is not None:
        B, H, W, _ = da.shape
        if not (mod.taps == 9 and 9 * cin <= DGRAD_TAP_SPLIT_MAX_N and B * H * W >= DGRAD_TAP_SPLIT_MIN_PIXELS):
            return False
    n = mod.conv.out_channels
    k = 3 if mod.taps == 9 else 1
    if not skip_wgrad:
        st.wgrad(x_act, cin, da, n, mod.taps, st.out(mod.conv.weight).view(n, mod.conv.weight.shape[1], k, k),
                 perm)   # staging order -> weight order
    if dgrad_out is not None:
        B, H, W, _ = da.shape
        if (mod.taps == 9 and dgrad_out.dtype == torch.float32 and 9 * cin <= DGRAD_TAP_SPLIT_MAX_N
                and B * H * W >= DGRAD_TAP_SPLIT_MIN_PIXELS):
            # few input channels: one 1x1 GEMM with N = 9*cin reads the gradient once (instead of once per tap, with
            # 16/32-column MMAs), then the nine shifted planes are summed
            wd9, cp, r8 = mod.packed_dgrad_taps(key, perm)
            planes = torch.empty(B, H, W, ops.pad_to(9 * r8, 64), device=da.device, dtype=torch.bfloat16)
            ops.conv_gemm(da, cp, wd9, 9 * r8, 1, None, None, "none", planes)
            ops.taps_gather_nhwc(planes, cin, r8, dgrad_out)
        else:
            wd, cp = mod.packed_dgrad(key, perm)
            ops.conv_gemm(da, cp, wd, cin, mod.taps, None, None, "none", dgrad_out)
    if acc is not None:
        wd9, cp, r8 = mod.packed_dgrad_taps(key, perm)
        planes = torch.empty(B, H, W, ops.pad_to(9 * r8, 64), device=da.device, dtype=torch.bfloat16)
        ops.conv_gemm(da, cp, wd9, 9 * r8, 1, None, None, "none", planes)
        ops.taps_gather_nhwc_acc(planes, cin, r8, acc[0], acc[1], acc[2])
        return True


def _norm_act_bwd(st, mod, dh, h, act_fn, keep=None):
    """Backward of h = act(ActNorm(conv)) for a Conv2dNorm: returns da (bf16 NHWC), accumulates d logs / d bias."""
    if mod.norm == "batchnorm":
        if not keep:
            raise NotImplementedError("recurrent-flows-msc_b200: eval-mode batch norm inside a training step is not supported")
        return _batchnorm_act_bwd(st, mod, dh, h, act_fn, keep)
    assert dh.shape[-1] == h.shape[-1]
    n = mod.conv.out_channels
    scale, _ = mod.norm_type.affine()
    da, _, _ = ops.act_affine_bwd(dh, h, n, scale, act_fn, 1.0, True, out_dv=st.out(mod.norm_type.bias),
                                  out_dvv=st.out(mod.norm_type.logs))
    return da


def _zeros_out_bwd(st, mod, dout, out):
    """Backward of a Conv2dZeros' output affine out = (conv + bias) * exp(3 logs) from fp32 NCHW (dout, out)."""
    B, n, H, W = out.shape
    dh, h = _nhwc(B, H, W, n, out.device), _nhwc(B, H, W, n, out.device)
    ops.pack_nhwc(dout, 0, n, dh, 0)
    ops.pack_nhwc(out, 0, n, h, 0)
    scale, _ = mod.affine()
    da, _, _ = ops.act_affine_bwd(dh, h, n, scale, "none", float(mod.logscale_factor), True, out_dv=st.out(mod.conv.bias),
                                  out_dvv=st.out(mod.logs))
    return da


# ----------------------------------------------------------------------------------------
# GlowStep
# ----------------------------------------------------------------------------------------
def _bnflow_fwd(step, x, ld):
    """flow_norm='batchnorm' (Flow/glow_modules.py:56-104) in a training step: batch statistics per POSITION over the batch
    dimension; the normalisation cannot be folded into the C x C mix, so it is its own elementwise kernel.  Returns the
    normalised tensor and what the backward needs."""
    bn = step.norm
    if bn.training:
        mean, var = ops.batch_stats_pos(x, bn.eps)            # [C,H,W]; var includes + eps
        with torch.no_grad():
            bn.running_mean.mul_(bn.momentum).add_(mean * (1 - bn.momentum))
            bn.running_var.mul_(bn.momentum).add_(var * (1 - bn.momentum))
    else:
        mean, var = bn.running_mean[0].float().contiguous(), bn.running_var[0].float().contiguous()
    a, c, d = bn._affine(mean, var, False)
    xn = ops.affine_pos(x, a, c)
    ld.add_(d)
    return xn, (mean, var, bn.training)


def _bnflow_bwd(st, step, x, dxn, saved):
    """Reverse mode of the batch-norm flow layer (torch ops on [B, C*H*W] views; a non-default option of the reference)."""
    bn = step.norm
    mean, var, batch_stats = saved
    lg, G = bn.log_gamma.detach()[0].float(), st.G.reshape(())
    sigma = var.sqrt()
    xc = x - mean
    xhat = xc / sigma
    gamma = torch.exp(lg)
    st.add(bn.beta, dxn.sum(0))
    st.add(bn.log_gamma, (dxn * xhat).sum(0) * gamma + G)
    dxhat = dxn * gamma
    if not batch_stats:                                        # running statistics: constants
        return (dxhat / sigma).contiguous()
    B = x.shape[0]
    dvar = (dxhat * xc).sum(0) * (-0.5) / (var * sigma) - 0.5 * G / var
    dmean = -(dxhat.sum(0)) / sigma
    return (dxhat / sigma + dvar * (2.0 / B) * xc + dmean / B).contiguous()


def _invconv_param_bwd(st, inv, dW, hw):
    """d loss / d InvConv parameters from d W and the log-det term, through the C x C assembly (torch autograd; the batched
    kernel covers the ActNorm-folded default, this the batch-norm variant)."""
    names = ("lower", "upper", "log_s") if inv.LU_decomposed else ("weight",)
    params = [getattr(inv, n) for n in names]
    with torch.enable_grad():
        leaf = [p.detach().float().requires_grad_() for p in params]
        if inv.LU_decomposed:
            l_mask, eye = inv._consts(leaf[0].device)
            lower = leaf[0] * l_mask + eye
            u = leaf[1] * l_mask.transpose(0, 1) + torch.diag(inv.sign_s * torch.exp(leaf[2]))
            Wm = torch.matmul(inv.p, torch.matmul(lower, u))
            ldw = leaf[2].sum()
        else:
            Wm = leaf[0]
            ldw = torch.linalg.slogdet(Wm)[1]
        gs = torch.autograd.grad([Wm, ldw * hw], leaf, [dW, st.G.reshape(())])
    for p, g in zip(params, gs):
        st.add(p, g)


def _glowstep_fwd(flow, step, x, ld, nn_template, cc, l, tape):
    aff = step.affine
    net = aff.net
    B, C, H, W = x.shape
    half, hid, act, dev = C // 2, aff.hidden_units, aff.non_lin, x.device
    if net[4].taps != 9 or 9 * C > TAP_SPLIT_MAX_N:
        raise NotImplementedError("recurrent-flows-msc_b200: coupling backward needs the tap-split form (C <= 256)")
    nn_in = nn_template.clone()
    bn_saved = xn = None
    if isinstance(step.norm, BatchNormFlow):
        xn, bn_saved = _bnflow_fwd(step, x, ld)
        Wm, per_pixel = step.invconv.weight_fwd()
        y = ops.mix1x1(xn, Wm, None, side=nn_in, side_n=half, side_off=cc, logdet=ld,
                       addend=(per_pixel * (H * W)).reshape(1).contiguous(), alpha=1.0)
    else:
        step.norm.maybe_initialize(x)
        Wf, bf = step._folded_fwd(H * W)[:2]
        y = ops.mix1x1(x, Wf, bf, side=nn_in, side_n=half, side_off=cc, logdet=ld, addend=step._dlogdet(H * W), alpha=1.0)
    # Recompute mode (flow.recompute / RFK_RECOMPUTE=1): the hidden tensors are NOT kept.  The coupling leaves
    # z1 = y[:, :C/2] unchanged, so the network's input is part of the step's OUTPUT: the backward regenerates
    # [cond | z1] -> h1, h2, tap planes from y with the same kernels (bit-identical), two steps at a time at most.  Only the
    # flow tensors x, y (C channels) stay on the tape: 10.8 GB -> 2.1 GB at the 570-frame workload.  Needs ActNorms whose
    # statistics are settled (a data-dependent init or batch-norm statistics must not run twice).
    recompute = getattr(flow, "recompute", None)
    recompute = ((RECOMPUTE if recompute is None else recompute) and bn_saved is None and net[0].foldable()
                 and net[2].foldable())
    h1, h2, taps = _coupling_nn(aff, nn_in, C, keep=not recompute)
    ops.coupling_tail_taps(taps, y, *aff.tail_params(), ld, False)
    if recompute:
        tape.append(lambda st: _glowstep_bwd(st, flow, step, x, y, None, None, None, None, cc, l, nn_template=nn_template))
    else:
        tape.append(lambda st: _glowstep_bwd(st, flow, step, x, y, nn_in, h1, h2, taps, cc, l, xn, bn_saved))
    return y


def _coupling_nn(aff, nn_in, C, keep):
    """The coupling network on the staged input: returns (h1, h2, taps); h1 / h2 are None when they were not asked for
    (keep=False) and the one-kernel path, which never materialises them, applies."""
    net = aff.net
    B, H, W, _ = nn_in.shape
    hid, act, dev = aff.hidden_units, aff.non_lin, nn_in.device
    wgt9, cp = net[4].packed_taps()
    taps = torch.empty(B, 9 * C, H, W, device=dev, dtype=torch.float32)
    if (FUSE_NN_TRAIN and not ops.SPLIT and net[2].taps == 1 and 9 * C <= FUSE_NN_MAX_PLANES and hid % 64 == 0 and hid <= 256
            and net[0].foldable() and net[2].foldable() and ops.gemm_m_tiles(B, H, W) >= FUSE_NN_MIN_TILES):
        # one kernel for the three convolutions (csrc/coupling_nn.cu); h1 / h2 leave as side outputs for the backward and
        # are not read back by the forward
        h1, h2 = (_nhwc(B, H, W, hid, dev), _nhwc(B, H, W, hid, dev)) if keep else (None, None)
        w1f, cp1 = net[0].packed_folded("cz", aff._perm(dev))
        ops.coupling_nn_fused(nn_in, cp1, net[0].taps, w1f, hid, net[2].packed_folded()[0], act, wgt9, 9 * C, taps, h1, h2)
        return h1, h2, taps
    h1, h2 = _nhwc(B, H, W, hid, dev), _nhwc(B, H, W, hid, dev)
    net[0].fused(nn_in, h1, act, "cz", aff._perm(dev))
    net[2].fused(h1, h2, act)
    ops.conv_gemm(h2, cp, wgt9, 9 * C, 1, None, None, "none", taps)
    return h1, h2, taps


def _glowstep_bwd(st, flow, step, x, zo, nn_in, h1, h2, taps, cc, l, xn=None, bn_saved=None, nn_template=None):
    aff = step.affine
    net = aff.net
    B, C, H, W = zo.shape
    half, hid, act, dev = C // 2, aff.hidden_units, aff.non_lin, zo.device
    if taps is None:
        # recompute mode: regenerate the network's activations from the step's output.  At most two steps' worth is alive: the
        # side stream's weight gradients of the step before last must be done before their inputs are released.
        st.recomputed += 1
        if st.recomputed % 2 == 0:
            st.join()
        nn_in = nn_template.clone()
        ops.pack_nhwc(zo, 0, half, nn_in, cc)
        h1, h2, taps = _coupling_nn(aff, nn_in, C, keep=True)
    dz = st.dz
    scale, shift, clamp, cs, csh = aff.tail_params()
    last = net[4]
    if clamp == "realnvp":
        outs = (st.out(last.logs), st.out(last.conv.bias), st.out(aff.scale), st.out(aff.scale_shift))
    else:
        outs = (st.out(last.logs), st.out(last.conv.bias)) + tuple(ops._zeros(half, dev) for _ in range(2))
    dsum = ops.coupling_taps_bwd(taps, zo, dz, scale, shift, clamp, cs, csh, st.g_ld, float(last.logscale_factor), outs)[0]
    if not WGRAD_SIDE_STREAM or B * H * W > WGRAD_SIDE_MAX_PIXELS:
        # everything that reads dS is stream-ordered before the next step's pack: one persistent buffer per shape whose pad
        # columns stay zero (a fresh zero-filled buffer per step was a 37 MB memset at level 1)
        dS = ops.workspace(("bwd_dS", C), (B, H, W, ops.cin_pad(C)), dev)
    else:
        dS = _nhwc(B, H, W, C, dev)      # small level: the side-stream weight gradient may still be reading the previous one
    ops.pack_nhwc(dsum, 0, C, dS, 0)
    cin = half + cc
    dnn = None
    if (FUSE_ACT_BWD and hid % 64 == 0 and hid <= 512 and net[0].norm == "actnorm" and net[2].norm == "actnorm"
            and ops.cin_pad(hid) == hid):
        # dh2 / dh1 never exist in HBM: each data-gradient GEMM applies act' and the ActNorm scale of the layer below in its
        # epilogue; the ActNorm parameter gradients come from the layers' own weight gradients
        _conv_bwd(st, last, h2, hid, dS)                                   # weight gradient of the last conv
        da2, r2 = _dgrad_actbwd(st, last, dS, net[2], h2, act)
        _wgrad_actnorm(st, net[2], h1, hid, da2, r2)
        da1, r1 = _dgrad_actbwd(st, net[2], da2, net[0], h1, act)
        _wgrad_actnorm(st, net[0], nn_in, cin, da1, r1, perm=aff._perm(dev))
        # network-input gradient: on the large levels the gather of the tap-split data gradient ADDS straight into the
        # condition's gradient and into dz[:, :half] (no dnn tensor, no rfk_add_channels launches)
        if not (FUSE_GATHER_ACC and _conv_bwd(st, net[0], nn_in, cin, da1, perm=aff._perm(dev), key="cz", skip_wgrad=True,
                                              acc=(st.cond_acc(l, B, cc, H, W) if cc else None, cc, dz))):
            dnn = torch.empty(B, cin, H, W, device=dev, dtype=torch.float32)
            _conv_bwd(st, net[0], nn_in, cin, da1, perm=aff._perm(dev), dgrad_out=dnn, key="cz", skip_wgrad=True)
    else:
        dnn = torch.empty(B, cin, H, W, device=dev, dtype=torch.float32)
        dh2 = _nhwc(B, H, W, hid, dev)
        _conv_bwd(st, last, h2, hid, dS, dgrad_out=dh2)
        da2 = _norm_act_bwd(st, net[2], dh2, h2, act)
        dh1 = _nhwc(B, H, W, hid, dev)
        _conv_bwd(st, net[2], h1, hid, da2, dgrad_out=dh1)
        da1 = _norm_act_bwd(st, net[0], dh1, h1, act)
        _conv_bwd(st, net[0], nn_in, cin, da1, perm=aff._perm(dev), dgrad_out=dnn, key="cz")
    if dnn is not None:
        ops.add_channels(dz, 0, dnn, cc, half)          # dz[:, :half] += d z1 (the network's input after the condition)
        if cc:
            st.add_cond(l, dnn, cc)
    if bn_saved is not None:     # flow_norm='batchnorm': plain InvConv mix, then the per-position batch-norm layer
        Wm = step.invconv.weight_fwd()[0]
        dWm, _ = ops.mix1x1_wgrad(xn, dz)
        dxn = ops.mix1x1(dz, Wm.t().contiguous(), None)
        _invconv_param_bwd(st, step.invconv, dWm, H * W)
        st.dz = _bnflow_bwd(st, step, x, dxn, bn_saved)
        return
    # ActNorm folded into the 1x1 mix: y = Wf x + bf
    fold = step._folded_fwd(H * W)
    buf = st.fold_buf(flow, step, C)
    dWf, dbf = buf[:C * C].view(C, C), buf[C * C:C * C + C]
    st.side_run(B * H * W, lambda: ops.mix1x1_wgrad(x, dz, out=buf), (x, dz, buf))   # only the fold backward at the level's end reads it
    st.dz = ops.mix1x1(dz, fold[3], None)           # Wf^T
    st.folds.append((step, dWf, dbf, H * W, buf))   # chained to the parameters in one batch at the end of the sweep


def _fold_bwd_all(st, flow, level=None):
    """Chain (d Wf, d bf, d logdet) of every GlowStep back to ActNorm's (bias, logs) and InvConv's parameters
    (Flow/glow_modules.py:33-54, 167-205).  LU-parameterised steps whose forward fold is registered for batched refresh go
    through ONE launch of rfk_fold_backward_batched over a pointer table (rebuilt only when a pointer changed); the rest
    take the batched-autograd path below."""
    folds, st.folds = st.folds, []
    fast, slow = [], []
    for it in folds:
        step = it[0]
        ok = step.invconv.LU_decomposed and step.invconv.__dict__.get("_perm32") is not None
        (fast if ok else slow).append(it)
    if fast:
        rows, key = [], []
        for step, dWf, dbf, hw, buf in fast:
            inv, C = step.invconv, step.invconv.w_shape[0]
            fold = step._folded_fwd(hw)
            outs = [st.out(p) for p in (step.norm.bias, step.norm.logs, inv.lower, inv.upper, inv.log_s)]
            row = [step.norm.bias.data_ptr(), step.norm.logs.data_ptr(), inv.lower.data_ptr(), inv.upper.data_ptr(),
                   inv.log_s.data_ptr(), inv.sign_s.data_ptr(), inv.__dict__["_perm32"].data_ptr(), C, hw, fold[0].data_ptr(),
                   dWf.data_ptr(), dbf.data_ptr(), buf[C * C + C:].data_ptr()] + [o.data_ptr() for o in outs]
            rows.append(row + [0] * (24 - len(row)))
            key.extend(row)
        key = tuple(key)
        tables = flow.__dict__.setdefault("_fold_bwd_tables", {})
        cache = tables.get(level)
        if cache is None or cache[0] != key:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("recurrent-flows-msc_b200: the fold-backward table changed inside a CUDA-graph capture; run the "
                                   "training step eagerly once with the same optimizer before capturing")
            cache = (key, torch.tensor([w for r in rows for w in r], dtype=torch.int64).to(st.dz.device))
            tables[level] = cache
        ops.call("rfk_fold_backward_batched", cache[1].data_ptr(), len(rows), st.G.data_ptr(), ops._stream())
    if slow:
        _fold_bwd_legacy(st, slow)


def _fold_bwd_legacy(st, folds):
    """Chain (d Wf, d bf, d logdet) of every GlowStep back to ActNorm's (bias, logs) and InvConv's parameters
    (Flow/glow_modules.py:33-54, 167-205).  These are C x C tensors, differentiated by torch autograd; steps with the
    same channel count and parameterisation (one flow level) are stacked and go through ONE batched autograd call,
    otherwise the ~50 tiny launches per step would cost more GPU time than the level-1 convolutions."""
    groups = {}
    for item in folds:
        step = item[0]
        groups.setdefault((step.invconv.w_shape[0], step.invconv.LU_decomposed, item[3]), []).append(item)
    for (C, lu, hw), items in groups.items():
        steps = [it[0] for it in items]
        inv0 = steps[0].invconv
        names = ("lower", "upper", "log_s") if lu else ("weight",)
        plist = [[s.norm.bias for s in steps], [s.norm.logs for s in steps]] + [[getattr(s.invconv, n) for s in steps] for n in names]
        with torch.enable_grad():
            leaf = [torch.stack([p.detach().float().reshape(p.shape if p.dim() <= 2 else (-1,)) for p in ps]).requires_grad_()
                    for ps in plist]
            bias, logs = leaf[0], leaf[1]                      # [K, C]
            if lu:
                l_mask, eye = inv0._consts(leaf[2].device)
                sign_s = torch.stack([s.invconv.sign_s for s in steps])
                perm = torch.stack([s.invconv.p for s in steps])
                lower = leaf[2] * l_mask + eye
                u = leaf[3] * l_mask.transpose(0, 1) + torch.diag_embed(sign_s * torch.exp(leaf[4]))
                Wm = torch.matmul(perm, torch.matmul(lower, u))
                ldw = leaf[4].sum(dim=1)
            else:
                Wm = leaf[2]
                ldw = torch.linalg.slogdet(Wm)[1]
            Wf = Wm * torch.exp(logs)[:, None, :]
            bfv = torch.matmul(Wf, bias[:, :, None])[:, :, 0]
            dl = (ldw + logs.sum(dim=1)) * hw
            dWf = torch.stack([it[1] for it in items])
            dbf = torch.stack([it[2] for it in items])
            gs = torch.autograd.grad([Wf, bfv, dl], leaf, [dWf, dbf, st.G.reshape(()).expand(len(items))])
        for ps, g in zip(plist, gs):
            for k, p in enumerate(ps):
                st.add(p, g[k])


# ----------------------------------------------------------------------------------------
# Split2d
# ----------------------------------------------------------------------------------------
def _split_fwd(sp, z, ld, nn_template, l, tape):
    B, C, H, W = z.shape
    half, cc, dev = sp._half, sp._cond, z.device
    sp_in = _nhwc(B, H, W, half + cc, dev)
    perm, t1 = None, None
    if sp.make_conditional:
        t1 = _nhwc(B, H, W, cc, dev)
        sp.convcond[0].fused(nn_template, t1, "relu")     # the level's condition sits at channels [0, cc)
        sp.convcond[2].fused(t1, sp_in, "relu")
        perm = sp._perm(dev)[0]
    ops.pack_nhwc(z, 0, half, sp_in, cc)
    params = torch.empty(B, 2 * half, H, W, device=dev, dtype=torch.float32)
    sp.conv[0].fused(sp_in, params, "cz", perm)
    z1 = torch.empty(B, half, H, W, device=dev, dtype=torch.float32)
    ops.copy_channels(z, 0, z1, 0, half)
    ops.gauss_logp(z, half, params, half, ops.PAIR_CROSS, sp.clamp_function, ld)
    tape.append(lambda st: _split_bwd(st, sp, z, params, sp_in, t1, nn_template, perm, l))
    return z1


def _split_bwd(st, sp, z, params, sp_in, t1, cbuf, perm, l):
    B, C, H, W = z.shape
    half, cc, dev = sp._half, sp._cond, z.device
    dzf = torch.zeros(B, 2 * half, H, W, device=dev, dtype=torch.float32)
    ops.copy_channels(st.dz, 0, dzf, 0, half)
    dparams = ops.gauss_logp_bwd(z, half, half, params, ops.PAIR_CROSS, sp.clamp_function, st.g_ld, dzf)
    conv = sp.conv[0]
    da = _zeros_out_bwd(st, conv, dparams, params)
    cin = half + cc
    dsp = _nhwc(B, H, W, cin, dev) if cc else None
    _conv_bwd(st, conv, sp_in, cin, da, perm=perm, dgrad_out=dsp, key="cz")
    # the z1 rows once more as fp32 NCHW for the main gradient
    dz1 = torch.empty(B, half, H, W, device=dev, dtype=torch.float32)
    wd, cp = conv._cache.get(("wd", "z1"), (conv.conv.weight,),
                             lambda: ops.pack_dgrad_weight(conv.conv.weight, sp._perm(dev)[1]))
    ops.conv_gemm(da, cp, wd, half, conv.taps, None, None, "none", dz1)
    ops.add_channels(dzf, 0, dz1, 0, half)
    if sp.make_conditional:
        c0, c2 = sp.convcond[0], sp.convcond[2]
        da2 = _norm_act_bwd(st, c2, dsp, sp_in, "relu")
        dt1 = _nhwc(B, H, W, cc, dev)
        _conv_bwd(st, c2, t1, cc, da2, dgrad_out=dt1)
        da1 = _norm_act_bwd(st, c0, dt1, t1, "relu")
        dc = torch.empty(B, cc, H, W, device=dev, dtype=torch.float32)
        _conv_bwd(st, c0, cbuf, cc, da1, dgrad_out=dc)
        st.add_cond(l, dc, cc)
    st.dz = dzf


# ----------------------------------------------------------------------------------------
# prior
# ----------------------------------------------------------------------------------------
def _prior_fwd(flow, z, base_condition, obj, tape):
    n = z.shape[1]
    if not flow.learn_prior:
        ops.gauss_logp(z, 0, None, n, ops.PAIR_SPLIT, "exp", obj)
        tape.append(lambda st: ops.gauss_logp_bwd(z, 0, n, None, ops.PAIR_SPLIT, "exp", st.g_ld, st.dz))
        return
    bc = ops.f32c(base_condition)
    B, Cb, H, W = bc.shape
    dev = bc.device
    u1, u2, act = flow.n_units_prior, flow.n_units_prior // 2, flow.non_lin_glow
    a0, a1, a2 = _nhwc(B, H, W, Cb, dev), _nhwc(B, H, W, u1, dev), _nhwc(B, H, W, u2, dev)
    ops.pack_nhwc(bc, 0, Cb, a0, 0)
    k0, k2 = {}, {}
    flow.prior[0].fused(a0, a1, act, keep=k0)
    flow.prior[2].fused(a1, a2, act, keep=k2)
    params = torch.empty(B, 2 * n, H, W, device=dev, dtype=torch.float32)
    flow.prior[4].fused(a2, params)
    ops.gauss_logp(z, 0, params, n, ops.PAIR_SPLIT, "exp", obj)

    def bwd(st):
        dparams = ops.gauss_logp_bwd(z, 0, n, params, ops.PAIR_SPLIT, "exp", st.g_ld, st.dz)
        da = _zeros_out_bwd(st, flow.prior[4], dparams, params)
        dh2 = _nhwc(B, H, W, u2, dev)
        _conv_bwd(st, flow.prior[4], a2, u2, da, dgrad_out=dh2)
        da2 = _norm_act_bwd(st, flow.prior[2], dh2, a2, act, k2)
        dh1 = _nhwc(B, H, W, u1, dev)
        _conv_bwd(st, flow.prior[2], a1, u1, da2, dgrad_out=dh1)
        da1 = _norm_act_bwd(st, flow.prior[0], dh1, a1, act, k0)
        st.dbase = torch.empty(B, Cb, H, W, device=dev, dtype=torch.float32)
        _conv_bwd(st, flow.prior[0], a0, Cb, da1, dgrad_out=st.dbase)
    tape.append(bwd)


def _level_done(st, flow, l):
    """End of level l in the reverse sweep: every parameter gradient of the level (and, the first time, of the prior) is
    final once the side stream has joined and the level's ActNorm / InvConv folds are chained back; a data-parallel
    optimizer may start all-reducing them (FlatAdam.attach)."""
    st.join()
    _fold_bwd_all(st, flow, l)
    hook = flow.__dict__.get("_rfk_grad_hook")
    if hook is not None:
        if not st.prior_announced:
            st.prior_announced = True
            hook("prior")
        hook(l)
    st.dz = ops.squeeze2d(st.dz, True)


# ----------------------------------------------------------------------------------------
# ListGlow.log_prob
# ----------------------------------------------------------------------------------------
def _log_prob_fwd(flow, x, conds, base_condition, obj0):
    """Forward of ListGlow.f + prior with a tape.  Returns (z, obj[B], tape)."""
    tape = []
    z = ops.f32c(x)
    B, dev = z.shape[0], z.device
    obj = obj0.to(device=dev, dtype=torch.float32).clone().contiguous()
    l = 0
    template = None
    for mod in flow.glow_frame:
        if isinstance(mod, Squeeze2d):
            z = ops.squeeze2d(z, False)
            tape.append(lambda st, l=l: _level_done(st, flow, l))
            cond = ops.f32c(conds[l])
            assert cond.shape[2:4] == z.shape[2:4], "condition and x in affine needs to match"
            cc = cond.shape[1]
            # zero-filled ALWAYS: only the condition channels [0, cc) are written here (every GlowStep fills the z1 region of its
            # own clone), but Split2d's convcond reads cin_pad(cc) channels of the template itself -- stale bf16 bit patterns
            # there (NaN times a zero weight is NaN) would poison the log-density
            ld_t = ops.cin_pad(z.shape[1] // 2 + cc)
            template = torch.zeros(B, z.shape[2], z.shape[3], ld_t, device=dev, dtype=torch.bfloat16)
            ops.pack_nhwc(cond, 0, cc, template, 0)
        elif isinstance(mod, Split2d):
            z = _split_fwd(mod, z, obj, template, l, tape)
            l += 1
        else:
            z = _glowstep_fwd(flow, mod, z, obj, template, cc, l, tape)
    _prior_fwd(flow, z, base_condition, obj, tape)
    return z, obj, tape


class _LogProb(torch.autograd.Function):
    """(z, nll) = ListGlow.log_prob with gradients for x, the conditions, the base condition and every parameter."""

    @staticmethod
    def forward(ctx, flow, n_cond, obj0, x, base_condition, *rest):
        conds, params = list(rest[:n_cond]), rest[n_cond:]
        z, obj, tape = _log_prob_fwd(flow, x, conds, base_condition, obj0)
        ctx.flow, ctx.tape, ctx.n_cond, ctx.params = flow, tape, n_cond, params
        ctx.cond_shapes = [c.shape for c in conds]
        ctx.has_base = base_condition is not None
        z_out = z.clone()   # the tape keeps z itself
        return z_out, -obj

    @staticmethod
    def backward(ctx, dz, dnll):
        if ctx.tape is None:
            raise RuntimeError("recurrent-flows-msc_b200: log_prob's tape was already consumed (no retain_graph)")
        dz = ops.f32c(dz).clone()
        g_ld = (-ops.f32c(dnll)).contiguous()
        st = _State(dz, g_ld, ctx.n_cond)
        tape, ctx.tape = ctx.tape, None
        with ops.zero_arena(dz.device):
            while tape:
                tape.pop()(st)
            st.join()
            _fold_bwd_all(st, ctx.flow)
        dconds = [st.dcond[i] for i in range(ctx.n_cond)]
        # parameters whose .grad was accumulated into directly get None here (nothing left for autograd to add)
        pgrads = [None if id(p) in st.direct else st.grads.get(id(p)) for p in ctx.params]
        pgrads = [None if g is None else g.view(p.shape).to(p.dtype) for g, p in zip(pgrads, ctx.params)]
        return (None, None, None, st.dz, st.dbase if ctx.has_base else None, *dconds, *pgrads)


def log_prob_with_grad(flow, x, conds, base_condition, obj0):
    params = [p for p in flow.parameters() if p.requires_grad]
    return _LogProb.apply(flow, len(conds), obj0, x, base_condition, *conds, *params)
