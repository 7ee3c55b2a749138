"""GlowStep and the multi-scale conditional ListGlow on librfk's sm_100a kernels.

Mirrors ``Flow/glow.py`` of cdglissov/recurrent-flows-msc (class names, constructor arguments,
method signatures, ``state_dict`` keys).  One GlowStep forward is four kernel launches:

  1. rfk_mix1x1              ActNorm folded into the invertible 1x1 conv, + bf16 z1 side output, + the
                             parameter-only log-det term
  2. rfk_conv_gemm           conv3x3 -> ActNorm -> ReLU                         (tcgen05, bf16 NHWC out via TMA store)
  3. rfk_conv1x1_taps_fused  conv1x1 -> ActNorm -> ReLU -> tap-split conv3x3    (tcgen05, hidden tile kept in TMEM)
  4. rfk_coupling_tail_taps  tap gather -> Conv2dZeros scale -> cross split -> clamp -> affine -> per-sample log-det

(steps 3 falls back to two rfk_conv_gemm launches, and 3+4 to rfk_conv_gemm_coupling, for shapes the fused kernels do
not cover) with no host synchronisation (the reference issues ~186 ATen ops and three .item() syncs per step).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..Utils.modules import ActFun
from .. import derived
from .glow_modules import (ActNorm, AffineCoupling, BatchNormFlow, Conv2dNorm, Conv2dZeros, InvConv,  # noqa: F401
                           Split2d, Squeeze2d, _Ctx, _ld_begin, _ld_end, _require_no_grad, _Versioned)

import os

# Fuse a coupling's tap gather + tail with the 1x1 mix that follows it (rfk_coupling_taps_mix) from this channel count up.
# Measured on B200 (570 frames / 30-frame sampling): the fused kernel has one thread per (pixel, 8 outputs) doing up to
# 4 x 18 gather loads serially and LOSES to the two specialised kernels at every level (6.56 vs 6.17 ms per step,
# sampling 2.63 vs 2.27 ms), so it is off by default; the entry point stays (tested) for shapes where launches dominate.
FUSE_GATHER_MIX_MIN_C = int(os.environ.get("RFK_FUSE_GATHER_MIX_MIN_C", str(1 << 30)))


class GlowStep(nn.Module):
    """Flow/glow.py:10-41: norm -> invconv -> affine coupling (and the exact inverse)."""

    def __init__(self, x_size, condition_size, args):
        super().__init__()
        self.n_units_affine = args.n_units_affine
        self.non_lin_glow = args.non_lin_glow
        self.clamp_type = args.clamp_type
        b, c, h, w = x_size
        if args.flow_norm == 'batchnorm':
            self.norm = BatchNormFlow(x_size, momentum=args.flow_batchnorm_momentum)
        else:
            self.norm = ActNorm(c)
        self.invconv = InvConv(c, LU_decomposed=args.LU_decomposed)
        self.affine = AffineCoupling(x_size, condition_size, hidden_units=self.n_units_affine,
                                     non_lin=self.non_lin_glow, clamp_type=self.clamp_type)
        self._cache = _Versioned()

    def _folded(self):
        """ActNorm folded into the 1x1 conv, both directions, plus the parameter-only log-det.

        fwd: y = W((x+b)*e^logs)          = (W diag(e^logs)) x + (W diag(e^logs)) b
        rev: x = (W^-1 y)*e^-logs - b     = (diag(e^-logs) W^-1) y - b
        per-pixel log-det = sum(logs) + log|det W|   (Flow/glow_modules.py:43,196)
        """
        def build():
            W, W_inv, per_pixel = self.invconv.matrices()
            logs = self.norm.logs.detach().float().reshape(-1)
            bias = self.norm.bias.detach().float().reshape(-1)
            s = torch.exp(logs)
            Wf = (W * s[None, :]).contiguous()
            bf = torch.mv(Wf, bias).contiguous()
            Wr = (W_inv / s[:, None]).contiguous()
            br = (-bias).contiguous()
            return Wf, bf, Wr, br, (per_pixel + logs.sum()).reshape(())
        return self._cache.get("fold", (self.norm.bias, self.norm.logs) + self.invconv._params(), build)

    def _folded_fwd(self, hw):
        """Forward direction only (no matrix inverses: density evaluation, training), for maps of hw pixels:
        (Wf, bf, per-pixel log-det (0-dim), Wf^T, hw * per-pixel log-det [1]).  In the LU form the entry is registered for
        the batched in-place refresh after optimizer steps (derived.py, rfk_fold_prepare_batched)."""
        inv = self.invconv

        def build():
            W, per_pixel = inv.weight_fwd()
            logs = self.norm.logs.detach().float().reshape(-1)
            bias = self.norm.bias.detach().float().reshape(-1)
            Wf = (W * torch.exp(logs)[None, :]).contiguous()
            pp = (per_pixel + logs.sum()).reshape(1)
            ld = torch.cat([pp, pp * hw]).contiguous()
            return Wf, torch.mv(Wf, bias).contiguous(), ld[0], Wf.t().contiguous(), ld[1:2], ld

        def register(cache, key, params, slot):
            if not inv.LU_decomposed:
                return
            Wf, bf, _, WfT, _, ld = slot[1]
            perm = inv.__dict__.get("_perm32")
            if perm is None or perm.device != Wf.device:
                perm = inv.p.argmax(dim=1).to(torch.int32).contiguous()    # row i of P has its one at column perm[i]
                inv.__dict__["_perm32"] = perm
            ptrs = [self.norm.bias, self.norm.logs, inv.lower, inv.upper, inv.log_s, inv.sign_s]
            if not all(t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda for t in ptrs):
                return
            C = inv.w_shape[0]
            derived.REFRESHER.register("fold", cache, key, params, slot,
                                       [t.data_ptr() for t in ptrs] + [perm.data_ptr(), C, hw, Wf.data_ptr(), WfT.data_ptr(),
                                                                       bf.data_ptr(), ld.data_ptr()], (Wf, bf, WfT, ld, perm))
        return self._cache.get(("fold_f", hw), (self.norm.bias, self.norm.logs) + inv._params(), build, register)

    def _dlogdet(self, hw):
        """H*W*(sum logs + log|det W|) as a cached device scalar (Flow/glow_modules.py:43,196)."""
        return self._folded_fwd(hw)[4]

    def forward(self, x, condition, logdet, reverse, _ctx=None):
        _require_no_grad()
        x = ops.f32c(x)
        B, C, H, W = x.shape
        own_ctx = _ctx is None
        if own_ctx:
            cc = condition.shape[1]
            nn_in = ops.workspace(("cpl_in", C // 2 + cc), (B, H, W, ops.buf_ld(C // 2 + cc)), x.device)
            ops.pack_nhwc(ops.f32c(condition), 0, cc, nn_in, 0)
            _ctx = _Ctx(nn_in, cc)
        cc = _ctx.cond_channels
        ld, extra = _ld_begin(logdet, B, x.device, inplace=not own_ctx)
        if isinstance(self.norm, BatchNormFlow):
            # per-position normalisation cannot be folded into the CxC mix: elementwise kernel + plain invconv
            Wm, Wm_inv, per_pixel = self.invconv.matrices()
            dl = self._cache.get(("dl_inv", H * W), self.invconv._params(), lambda: (per_pixel * (H * W)).reshape(1).contiguous())
            if not reverse:
                x, ld = self.norm(x, ld, False)
                y = ops.mix1x1(x, Wm, None, side=_ctx.nn_in, side_n=C // 2, side_off=cc, logdet=ld,
                               addend=None if ld is None else dl, alpha=1.0)
                _ctx.z1_packed = True
                y, ld = self.affine(y, condition, ld, False, _ctx=_ctx)
                _ctx.z1_packed = False
                return y, _ld_end(ld, extra)
            if own_ctx:
                x = x.clone()
            y, ld = self.affine(x, condition, ld, True, _ctx=_ctx)
            y = ops.mix1x1(y, Wm_inv, None, logdet=ld, addend=None if ld is None else dl, alpha=-1.0)
            out, ld = self.norm(y, ld, True)
            _ctx.z1_packed = False   # the next reverse step packs z1 itself (the side output would be pre-normalisation)
            return out, _ld_end(ld, extra)
        fuse_gm = C >= FUSE_GATHER_MIX_MIN_C
        if not reverse:
            pend = _ctx.pending
            if pend is not None and pend[2].data_ptr() == x.data_ptr() and self.norm.is_initialized():
                # the previous step's coupling tail (tap gather + affine) is still pending: absorb it into this step's mix
                _ctx.pending = None
                Wf, bf = self._folded_fwd(H * W)[:2]
                dl = None if ld is None else self._dlogdet(H * W)
                y = ops.coupling_taps_mix(pend[1], x, *pend[0].tail_params(), pend[3], False, Wf, bf,
                                          side=_ctx.nn_in, side_n=C // 2, side_off=cc, logdet=ld, addend=dl, alpha=1.0)
            else:
                _ctx.flush()
                self.norm.maybe_initialize(x)
                Wf, bf = self._folded_fwd(H * W)[:2]
                dl = None if ld is None else self._dlogdet(H * W)
                y = ops.mix1x1(x, Wf, bf, side=_ctx.nn_in, side_n=C // 2, side_off=cc, logdet=ld, addend=dl, alpha=1.0)
            _ctx.z1_packed = True
            kind, t = self.affine.run_nn(y, condition, _ctx)
            _ctx.z1_packed = False
            if kind == "taps" and not own_ctx and fuse_gm:
                _ctx.pending = (self.affine, t, y, ld)    # the next module's mix (or ListGlow's flush) applies the tail
            else:
                self.affine.finish(kind, t, y, ld, False)
            return y, _ld_end(ld, extra)
        kind, t = self.affine.run_nn(x, condition, _ctx)
        if kind == "taps" and self.norm.is_initialized() and fuse_gm:
            # coupling inverse (tap gather) + InvConv^-1 + ActNorm^-1 in one launch; x itself is not modified
            _, _, Wr, br, _ = self._folded()
            dl = None if ld is None else self._dlogdet(H * W)
            out = ops.coupling_taps_mix(t, x, *self.affine.tail_params(), ld, True, Wr, br, side=_ctx.nn_in,
                                        side_n=C // 2, side_off=cc, logdet=ld, addend=dl, alpha=-1.0)
        else:
            y = x.clone() if own_ctx else x   # the coupling inverse works in place
            self.affine.finish(kind, t, y, ld, True)
            if not self.norm.is_initialized():
                # reference order (Flow/glow.py:37-40): affine^-1, invconv^-1, THEN the ActNorm -- a fresh training-mode
                # ActNorm takes its data-dependent statistics from the invconv^-1 output
                _, W_inv, per_pixel = self.invconv.matrices()
                y = ops.mix1x1(y, W_inv, None)
                self.norm.maybe_initialize(y)
                out = ops.actnorm(y, self.norm.bias.data, self.norm.logs.data, True)
                if ld is not None:
                    ld -= self._dlogdet(H * W)
                _ctx.z1_packed = False
                return out, _ld_end(ld, extra)
            _, _, Wr, br, _ = self._folded()
            dl = None if ld is None else self._dlogdet(H * W)
            out = ops.mix1x1(y, Wr, br, side=_ctx.nn_in, side_n=C // 2, side_off=cc, logdet=ld, addend=dl, alpha=-1.0)
        _ctx.z1_packed = True   # the next reverse step of this level reads this z1
        return out, _ld_end(ld, extra)


class ListGlow(nn.Module):
    """Flow/glow.py:43-160: L levels of (Squeeze2d, K x GlowStep, Split2d) with a learned or N(0,1) prior."""

    def __init__(self, x_size, condition_size, base_dist_size, args):
        super().__init__()
        assert isinstance(condition_size, list), "condition_size is not a list, make sure it fits L"
        # training memory mode (not a parameter, not in the state_dict): True = regenerate the coupling networks' activations
        # from each GlowStep's output in the backward instead of keeping them (Flow/training.py), None = RFK_RECOMPUTE
        self.recompute = None
        self.learn_prior = args.learn_prior
        self.n_units_prior = args.n_units_prior
        self.make_conditional = args.make_conditional
        self.base_norm = args.base_norm
        self.non_lin_glow = args.non_lin_glow
        self.conditional_clamp_function = args.split2d_act
        self.L = args.L
        self.K = args.K
        self.n_bits = args.n_bits
        Bx, Cx, Hx, Wx = x_size
        Bc, Cc, Hc, Wc = base_dist_size
        layers = []
        for l in range(0, self.L):
            layers.append(Squeeze2d())
            Cx, Hx, Wx = Cx * 4, Hx // 2, Wx // 2
            x_size = [Bx, Cx, Hx, Wx]
            condition_size_cur = condition_size[l]
            for i in range(0, self.K):
                layers.append(GlowStep(x_size, condition_size_cur, args))
            if l < (self.L - 1):
                layers.append(Split2d(x_size, condition_size_cur, self.make_conditional,
                                      self.conditional_clamp_function))
                Cx = Cx // 2
                x_size = [Bx, Cx, Hx, Wx]
        self.glow_frame = nn.ModuleList(layers)
        self._z_channels = Cx
        if self.learn_prior:
            self.prior = nn.Sequential(
                Conv2dNorm(Cc, self.n_units_prior, norm=self.base_norm),
                ActFun(self.non_lin_glow),
                Conv2dNorm(self.n_units_prior, self.n_units_prior // 2, norm=self.base_norm),
                ActFun(self.non_lin_glow),
                Conv2dZeros(in_channel=self.n_units_prior // 2, out_channel=2 * Cx),
            )
            self._base_channels = Cc
        # learn_prior == False: the reference's zero `prior_in` is mean = log_scale = 0, which the
        # Gaussian kernels take as a null parameter pointer.

    # -- level bookkeeping ------------------------------------------------------------------
    def _level_ctx(self, z, cond):
        B, C, H, W = z.shape
        cc = cond.shape[1]
        assert cond.shape[2:4] == z.shape[2:4], "condition and x in affine needs to match"
        nn_in = ops.workspace(("lvl_in", C // 2 + cc), (B, H, W, ops.buf_ld(C // 2 + cc)), z.device)
        ops.pack_nhwc(ops.f32c(cond), 0, cc, nn_in, 0)
        return _Ctx(nn_in, cc)

    def g(self, z, condition, logdet, temperature, eps_list=None):
        """z -> x (Flow/glow.py:90-102).  ``eps_list[l]`` optionally injects level l's Split2d draw."""
        _require_no_grad()
        x = ops.f32c(z).clone()
        l = len(condition) - 1
        ld, extra = _ld_begin(logdet, x.shape[0], x.device)
        ctx = None
        for step in reversed(self.glow_frame):
            if isinstance(step, Squeeze2d):
                x = step(x, undo_squeeze=True)
                ctx = None
            elif isinstance(step, Split2d):
                l = l - 1
                ctx = self._level_ctx_for_split(x, condition[l], step)
                x, ld = step(x, condition[l], logdet=ld, reverse=True, temperature=temperature, _ctx=ctx,
                             eps=None if eps_list is None else eps_list[l])
            else:
                if ctx is None:
                    ctx = self._level_ctx(x, condition[l])
                x, ld = step(x, condition[l], logdet=ld, reverse=True, _ctx=ctx)
        return x, _ld_end(ld, extra)

    def _level_ctx_for_split(self, z1, cond, split):
        # reverse direction: the level's tensor has 2*half channels once Split2d has re-attached z2
        B, half, H, W = z1.shape
        cc = cond.shape[1]
        nn_in = ops.workspace(("lvl_in", half + cc), (B, H, W, ops.buf_ld(half + cc)), z1.device)
        ops.pack_nhwc(ops.f32c(cond), 0, cc, nn_in, 0)
        return _Ctx(nn_in, cc)

    def f(self, x, condition, logdet):
        """x -> z (Flow/glow.py:105-117)."""
        _require_no_grad()
        z = ops.f32c(x)
        l = 0
        ld, extra = _ld_begin(logdet, z.shape[0], z.device)
        ctx = None
        for step in self.glow_frame:
            if ctx is not None and not isinstance(step, GlowStep):
                ctx.flush()   # a coupling tail left pending for a following mix: no mix follows, apply it now
            if isinstance(step, Squeeze2d):
                z = step(z, undo_squeeze=False)
                ctx = self._level_ctx(z, condition[l])
            elif isinstance(step, Split2d):
                z, ld = step(z, condition[l], logdet=ld, reverse=False, _ctx=ctx)
                l = l + 1
            else:
                z, ld = step(z, condition[l], logdet=ld, reverse=False, _ctx=ctx)
        if ctx is not None:
            ctx.flush()
        return z, _ld_end(ld, extra)

    def uniform_binning_correction(self, x):
        """Flow/glow.py:119-126 (the U(0, 2^-n_bits) draw stays in torch's generator)."""
        n_bins = 2 ** self.n_bits
        b, c, h, w = x.size()
        x_noise = x + torch.zeros_like(x).uniform_(0, 1.0 / n_bins)
        objective = -np.log(n_bins) * (c * h * w) * torch.ones(b, device=x.device)
        return x_noise, objective

    def _prior_params(self, base_condition):
        if not self.learn_prior:
            return None
        bc = ops.f32c(base_condition)
        B, Cb, H, W = bc.shape
        dev = bc.device
        u1, u2 = self.n_units_prior, self.n_units_prior // 2
        a0 = ops.workspace(("pr_in", Cb), (B, H, W, ops.buf_ld(Cb)), dev)
        a1 = ops.workspace(("pr_h1", u1), (B, H, W, ops.buf_ld(u1)), dev)
        a2 = ops.workspace(("pr_h2", u2), (B, H, W, ops.buf_ld(u2)), dev)
        ops.pack_nhwc(bc, 0, Cb, a0, 0)
        self.prior[0].fused(a0, a1, self.non_lin_glow)
        self.prior[2].fused(a1, a2, self.non_lin_glow)
        params = torch.empty(B, 2 * self._z_channels, H, W, device=dev, dtype=torch.float32)
        return self.prior[4].fused(a2, params)

    def _wants_grad(self, x, condition, base_condition):
        """The tape-recording path is taken only when something could receive a gradient: a parameter of a training-mode
        flow, or an input that requires grad.  An eval-mode call outside torch.no_grad() (the reference's evaluation code)
        whose inputs need no gradient runs the inference kernels."""
        if any(torch.is_tensor(t) and t.requires_grad for t in [x, base_condition] + list(condition)):
            return True
        return self.training and any(p.requires_grad for p in self.parameters())

    def log_prob(self, x, condition, base_condition, logdet=0, noise=None):
        """Flow/glow.py:128-141.  Returns (z, nll[B]).  ``noise`` optionally injects the dequantisation draw."""
        assert isinstance(condition, list), "Condition is not a list, make sure it fits L"
        if noise is None:
            x, obj_unif = self.uniform_binning_correction(x)
        else:
            b, c, h, w = x.size()
            obj_unif = -np.log(2 ** self.n_bits) * (c * h * w) * torch.ones(b, device=x.device)
            x = x + noise
        if torch.is_grad_enabled() and self._wants_grad(x, condition, base_condition):
            # training: tape-recording forward + hand-written backward kernels (Flow/training.py)
            if ops.SPLIT:
                raise NotImplementedError("recurrent-flows-msc_b200: RFK_CONV_PRECISION=bf16x3 covers density evaluation and sampling; "
                                          "train in the default bf16 mode")
            from .training import log_prob_with_grad
            if torch.is_tensor(logdet) and logdet.requires_grad:
                raise NotImplementedError("recurrent-flows-msc_b200: gradients w.r.t. the logdet argument are not implemented")
            obj0 = obj_unif + (logdet.detach() if torch.is_tensor(logdet) else logdet)
            return log_prob_with_grad(self, x, condition, base_condition, obj0)
        with torch.no_grad():
            z, obj = self.f(x, condition, logdet)
            if not torch.is_tensor(obj) or obj.dim() != 1:
                obj = torch.zeros(z.shape[0], device=z.device) + obj
            obj = obj + obj_unif
            ops.gauss_logp(z, 0, self._prior_params(base_condition), z.shape[1], ops.PAIR_SPLIT, "exp", obj)
            return z, -obj

    def sample(self, z, condition, base_condition, num_samples=32, temperature=0.8, eval_params=False,
               eps_prior=None, eps_list=None):
        """Flow/glow.py:143-160."""
        with torch.no_grad():
            mean = std = None
            if z is None:
                params = self._prior_params(base_condition)
                ref = ops.f32c(condition[-1])
                hh, ww = ref.shape[2], ref.shape[3]
                n = params.shape[0] if params is not None else num_samples
                cz = self._z_channels
                if eps_prior is None:
                    eps_prior = torch.randn(n, cz, hh, ww, device=ref.device, dtype=torch.float32)
                z = torch.empty(n, cz, hh, ww, device=ref.device, dtype=torch.float32)
                ops.gauss_sample(ops.f32c(eps_prior), params, cz, ops.PAIR_SPLIT, "exp", temperature, z, 0)
                if eval_params:
                    if params is None:
                        mean, std = torch.zeros_like(z), torch.ones_like(z)
                    else:
                        mean, std = params[:, :cz], torch.exp(params[:, cz:])
            # the reverse pass is a chain of ~280 small dependent launches: programmatic dependent launch lets every kernel's
            # prologue overlap its predecessor's tail (-6 % per frame; the throughput-bound passes do not use it)
            with ops.pdl(True):
                x, _ = self.g(z, condition, logdet=None, temperature=temperature, eps_list=eps_list)
        if eval_params:
            return x, (mean, std)
        return x
