"""Generate golden fixtures from the reference's own Python modules.

Run in the build container only (it imports /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

Each fixture is a ``torch.save``d dict of plain tensors / python scalars:
inputs, the reference module's ``state_dict`` and the reference outputs.  The
fixtures are what pins ``oracle/`` (tests/test_oracle_golden.py) and what the
GPU parity tests replay through the CUDA path.  The reference has no tests or
golden vectors of its own for this path.
"""
import os
import sys
import types
import warnings

import torch
import torch.distributions as td

REF = os.environ.get("RFMSC_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from Flow import ListGlow  # noqa: E402
from Flow.glow import GlowStep  # noqa: E402
from Flow.glow_modules import ActNorm, AffineCoupling, BatchNormFlow, InvConv, Split2d, Squeeze2d  # noqa: E402
from Utils import ConvLSTM  # noqa: E402


def save(name, d):
    path = os.path.join(OUT, name + ".pt")
    torch.save(d, path)
    print(f"{name:32s} {os.path.getsize(path) / 1024:8.1f} KiB")


def sd_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def perturb(m, gen, scale=0.15):
    """Make a freshly built module 'trained-like': zero-initialised tensors
    (Conv2dZeros, realnvp scale, ActNorm bias/logs) become non-trivial and every
    ActNorm is marked initialised, so no data-dependent init runs."""
    with torch.no_grad():
        for name, p in m.named_parameters():
            p.add_(torch.randn(p.shape, generator=gen) * scale * (0.3 if "conv.weight" in name else 1.0))
        for name, b in m.named_buffers():
            if name.endswith("initialized"):
                b.fill_(1)


class FixedNormalSample:
    """Replace td.Normal.sample by loc + scale*eps with eps popped from a queue."""

    def __init__(self, eps_list):
        self.eps = list(eps_list)

    def __enter__(self):
        self._orig = td.Normal.sample
        outer = self

        def sample(self_, sample_shape=torch.Size()):
            e = outer.eps.pop(0)
            return self_.loc + self_.scale * e

        td.Normal.sample = sample
        return self

    def __exit__(self, *a):
        td.Normal.sample = self._orig


def main():
    g = torch.Generator().manual_seed(1234)
    R = lambda *s: torch.randn(*s, generator=g)  # noqa: E731

    # ---- squeeze -----------------------------------------------------------
    x = R(2, 3, 4, 6)
    sq = Squeeze2d()
    y = sq(x, undo_squeeze=False)
    save("squeeze", {"x": x, "y": y, "undo": sq(y, undo_squeeze=True)})

    # ---- actnorm -----------------------------------------------------------
    x = R(3, 5, 4, 4) * 2.0 + 0.7
    m = ActNorm(5).train()
    y, ld = m(x, logdet=torch.zeros(3), reverse=False)
    d = {"x": x, "y_init_train": y.detach(), "logdet_init_train": ld.detach(), "sd_after_init": sd_of(m)}
    xr, ldr = m(y.detach(), logdet=ld.detach(), reverse=True)
    d.update(x_rev=xr.detach(), logdet_rev=ldr.detach())
    m2 = ActNorm(5).eval()  # quirk: first call in eval marks initialised without initialising
    y2, _ = m2(x, logdet=None, reverse=False)
    d.update(y_eval_first=y2.detach(), sd_eval_first=sd_of(m2))
    y3, ld3 = m(x[:2] * 0.5, logdet=0.0, reverse=False)  # python-scalar logdet -> 0-dim tensor
    d.update(x3=x[:2] * 0.5, y3=y3.detach(), logdet3=ld3.detach())
    save("actnorm", d)

    # ---- invconv -----------------------------------------------------------
    for lu in (True, False):
        torch.manual_seed(7)
        m = InvConv(6, LU_decomposed=lu)
        perturb(m, g, 0.05)
        x = R(2, 6, 3, 5)
        z, ld = m(x, logdet=torch.zeros(2), reverse=False)
        xr, ldr = m(z.detach(), logdet=ld.detach(), reverse=True)
        save("invconv_lu" if lu else "invconv_plain",
             {"x": x, "sd": sd_of(m), "z": z.detach(), "logdet": ld.detach(),
              "x_rev": xr.detach(), "logdet_rev": ldr.detach()})

    # ---- affine coupling -----------------------------------------------------
    for clamp, non_lin in (("realnvp", "relu"), ("glow", "relu"), ("softclamp", "leakyrelu"), ("none", "relu")):
        torch.manual_seed(11)
        m = AffineCoupling([2, 8, 6, 6], [2, 5, 6, 6], hidden_units=16, non_lin=non_lin, clamp_type=clamp).eval()
        perturb(m, g)
        x, c = R(2, 8, 6, 6), R(2, 5, 6, 6)
        z, ld = m(x, c, logdet=torch.zeros(2), reverse=False)
        xr, ldr = m(z.detach(), c, logdet=ld.detach(), reverse=True)
        save(f"coupling_{clamp}", {"x": x, "cond": c, "sd": sd_of(m), "z": z.detach(), "logdet": ld.detach(),
                                   "x_rev": xr.detach(), "logdet_rev": ldr.detach(),
                                   "clamp_type": clamp, "non_lin": non_lin, "hidden": 16})

    # ---- split2d -----------------------------------------------------------
    for cond, clampf in ((True, "softplus"), (False, "exp")):
        torch.manual_seed(13)
        m = Split2d([2, 8, 4, 4], [2, 6, 4, 4], make_conditional=cond, clamp_function=clampf).eval()
        perturb(m, g)
        x, c = R(2, 8, 4, 4), R(2, 6, 4, 4)
        z1, ld = m(x, c, logdet=torch.zeros(2), reverse=False)
        eps = R(2, 4, 4, 4)
        with FixedNormalSample([eps]):
            xr, _ = m(z1.detach(), c, logdet=None, reverse=True, temperature=0.7)
        save(f"split2d_{'cond' if cond else 'uncond'}_{clampf}",
             {"x": x, "cond": c, "sd": sd_of(m), "z1": z1.detach(), "logdet": ld.detach(), "eps": eps,
              "temperature": 0.7, "x_rev": xr.detach(), "make_conditional": cond, "clamp_function": clampf})

    # ---- glow step ---------------------------------------------------------
    args = types.SimpleNamespace(LU_decomposed=True, n_units_affine=16, non_lin_glow="relu", clamp_type="realnvp",
                                 flow_norm="actnorm", flow_batchnorm_momentum=0.0)
    torch.manual_seed(17)
    m = GlowStep([2, 8, 4, 4], [2, 3, 4, 4], args).eval()
    perturb(m, g)
    x, c = R(2, 8, 4, 4), R(2, 3, 4, 4)
    z, ld = m(x, c, logdet=torch.zeros(2), reverse=False)
    xr, ldr = m(z.detach(), c, logdet=ld.detach(), reverse=True)
    save("glowstep", {"x": x, "cond": c, "sd": sd_of(m), "z": z.detach(), "logdet": ld.detach(),
                      "x_rev": xr.detach(), "logdet_rev": ldr.detach(), "hidden": 16})

    # ---- ListGlow, conditional with learned prior ----------------------------
    def glow_args(**kw):
        a = dict(LU_decomposed=True, n_units_affine=16, non_lin_glow="relu", clamp_type="realnvp",
                 flow_norm="actnorm", flow_batchnorm_momentum=0.0, learn_prior=True, n_units_prior=16,
                 make_conditional=True, base_norm="actnorm", split2d_act="softplus", L=3, K=2, n_bits=8)
        a.update(kw)
        return types.SimpleNamespace(**a)

    B = 2
    a = glow_args()
    cond_sizes = [[B, 4, 8, 8], [B, 6, 4, 4], [B, 8, 2, 2]]
    torch.manual_seed(19)
    m = ListGlow([B, 1, 16, 16], cond_sizes, [B, 5, 2, 2], a).eval()
    perturb(m, g)
    x = torch.floor(torch.rand(B, 1, 16, 16, generator=g) * 256) / 256 - 0.5
    conds = [R(*s) for s in cond_sizes]
    base = R(B, 5, 2, 2)
    z, ld = m.f(x, conds, logdet=0.0)
    torch.manual_seed(23)
    noise = torch.zeros_like(x).uniform_(0, 1.0 / 2 ** a.n_bits)
    torch.manual_seed(23)
    z_lp, nll = m.log_prob(x, conds, base, logdet=0)
    eps_prior = R(B, 16, 2, 2)
    eps_split = [R(B, 2, 8, 8), R(B, 4, 4, 4)]  # level 0, level 1 (consumed in reverse order: level 1 first)
    with FixedNormalSample([eps_prior, eps_split[1], eps_split[0]]):
        xs = m.sample(None, conds, base, num_samples=B, temperature=0.8)
    save("listglow_cond", {"x": x, "cond": conds, "base": base, "sd": sd_of(m), "z_f": z.detach(),
                           "logdet_f": ld.detach(), "noise": noise, "z_logprob": z_lp.detach(), "nll": nll.detach(),
                           "eps_prior": eps_prior, "eps_split": eps_split, "x_sample": xs.detach(),
                           "temperature": 0.8, "args": vars(a), "x_size": [B, 1, 16, 16],
                           "cond_sizes": cond_sizes, "base_size": [B, 5, 2, 2],
                           "bpd": (nll / (torch.log(torch.tensor(2.0)) * 256)).detach()})

    # ---- ListGlow, cfg1-shaped but narrow: unconditional, fixed N(0,1) prior -----
    a = glow_args(learn_prior=False, make_conditional=False, L=3, K=2, clamp_type="glow")
    cond_sizes = [[B, 0, 8, 8], [B, 0, 4, 4], [B, 0, 2, 2]]
    torch.manual_seed(29)
    m = ListGlow([B, 1, 16, 16], cond_sizes, [B, 0, 2, 2], a).eval()
    perturb(m, g)
    conds = [torch.zeros(*s) for s in cond_sizes]
    torch.manual_seed(31)
    noise = torch.zeros_like(x).uniform_(0, 1.0 / 2 ** a.n_bits)
    torch.manual_seed(31)
    z_lp, nll = m.log_prob(x, conds, None, logdet=0)
    eps_prior = R(B, 16, 2, 2)
    eps_split = [R(B, 2, 8, 8), R(B, 4, 4, 4)]
    with FixedNormalSample([eps_prior, eps_split[1], eps_split[0]]):
        xs = m.sample(None, conds, None, num_samples=B, temperature=0.9)
    save("listglow_uncond", {"x": x, "sd": sd_of(m), "noise": noise, "z_logprob": z_lp.detach(), "nll": nll.detach(),
                             "eps_prior": eps_prior, "eps_split": eps_split, "x_sample": xs.detach(),
                             "temperature": 0.9, "args": vars(a), "x_size": [B, 1, 16, 16],
                             "cond_sizes": cond_sizes, "base_size": [B, 0, 2, 2]})

    # ---- BatchNormFlow (flow_norm='batchnorm') and a GlowStep / ListGlow built on it, base_norm='batchnorm' ----------
    torch.manual_seed(41)
    m = BatchNormFlow([4, 3, 4, 5], momentum=0.25).train()
    perturb(m, g, 0.3)
    x = R(4, 3, 4, 5) * 1.5 + 0.2
    y, ld = m(x, logdet=torch.zeros(4), reverse=False)        # training: batch statistics, running buffers updated
    d = {"x": x, "sd0": {"log_gamma": m.log_gamma.detach().clone(), "beta": m.beta.detach().clone()},
         "momentum": 0.25, "y_train": y.detach(), "logdet_train": ld.detach(), "sd_after": sd_of(m)}
    m.eval()
    y2, ld2 = m(x[:2] * 0.7, logdet=torch.zeros(2), reverse=False)
    xr, ldr = m(y2.detach(), logdet=ld2.detach(), reverse=True)
    d.update(x2=x[:2] * 0.7, y_eval=y2.detach(), logdet_eval=ld2.detach(), x_rev=xr.detach(), logdet_rev=ldr.detach())
    save("batchnormflow", d)

    a = glow_args(flow_norm="batchnorm", base_norm="batchnorm", flow_batchnorm_momentum=0.0, L=2, K=2)
    cond_sizes = [[B, 4, 8, 8], [B, 6, 4, 4]]
    torch.manual_seed(43)
    m = ListGlow([B, 1, 16, 16], cond_sizes, [B, 5, 4, 4], a).eval()
    perturb(m, g)
    with torch.no_grad():
        for name, buf in m.named_buffers():
            if name.endswith("running_var"):
                buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)
            elif name.endswith("running_mean"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.2)
    x = torch.floor(torch.rand(B, 1, 16, 16, generator=g) * 256) / 256 - 0.5
    conds = [R(*s) for s in cond_sizes]
    base = R(B, 5, 4, 4)
    torch.manual_seed(47)
    noise = torch.zeros_like(x).uniform_(0, 1.0 / 2 ** a.n_bits)
    torch.manual_seed(47)
    z_lp, nll = m.log_prob(x, conds, base, logdet=0)
    eps_prior = R(B, 8, 4, 4)
    eps_split = [R(B, 2, 8, 8)]
    with FixedNormalSample([eps_prior, eps_split[0]]):
        xs = m.sample(None, conds, base, num_samples=B, temperature=0.8)
    save("listglow_batchnorm", {"x": x, "cond": conds, "base": base, "sd": sd_of(m), "noise": noise,
                                "z_logprob": z_lp.detach(), "nll": nll.detach(), "eps_prior": eps_prior,
                                "eps_split": eps_split, "x_sample": xs.detach(), "temperature": 0.8, "args": vars(a),
                                "x_size": [B, 1, 16, 16], "cond_sizes": cond_sizes, "base_size": [B, 5, 4, 4]})

    # ---- ConvLSTM ------------------------------------------------------------
    torch.manual_seed(37)
    m = ConvLSTM(in_channels=3, hidden_channels=4, kernel_size=[3, 3], bias=True, peephole=True)
    x = R(2, 3, 3, 5, 6)
    out, h, c = m(x)
    h0, c0 = R(2, 4, 5, 6), R(2, 4, 5, 6)
    out2, h2, c2 = m(x[:, :1], h0, c0)
    sd = {k: v for k, v in sd_of(m).items() if "conv" in k}  # peepholes are zeros (Utils/modules.py:385-389)
    save("convlstm", {"x": x, "sd": sd, "out": out.detach(), "h": h.detach(), "c": c.detach(),
                      "h0": h0, "c0": c0, "out2": out2.detach(), "h2": h2.detach(), "c2": c2.detach()})

    # training mode of the same options: batch statistics in BatchNormFlow (per position) and in the prior's BatchNorm2d,
    # with the gradients torch autograd gives through the reference's modules (pins the oracle's training flag)
    B4 = 4
    a = glow_args(flow_norm="batchnorm", base_norm="batchnorm", flow_batchnorm_momentum=0.0, L=2, K=2)
    cond_sizes = [[B4, 4, 8, 8], [B4, 6, 4, 4]]
    torch.manual_seed(53)
    m = ListGlow([B4, 1, 16, 16], cond_sizes, [B4, 5, 4, 4], a).train()
    perturb(m, g)
    sd0 = sd_of(m)
    x = torch.floor(torch.rand(B4, 1, 16, 16, generator=g) * 256) / 256 - 0.5
    conds = [R(*s) for s in cond_sizes]
    base = R(B4, 5, 4, 4)
    torch.manual_seed(59)
    noise = torch.zeros_like(x).uniform_(0, 1.0 / 2 ** a.n_bits)
    torch.manual_seed(59)
    z_lp, nll = m.log_prob(x, conds, base, logdet=0)
    wts = torch.rand(B4, generator=g) + 0.5
    ((nll * wts).sum() / (0.6931471805599453 * 256 * B4)).backward()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    save("listglow_batchnorm_train", {"x": x, "cond": conds, "base": base, "sd": sd0, "noise": noise, "wts": wts,
                                      "z_logprob": z_lp.detach(), "nll": nll.detach(), "grads": grads, "args": vars(a),
                                      "x_size": [B4, 1, 16, 16], "cond_sizes": cond_sizes, "base_size": [B4, 5, 4, 4]})


if __name__ == "__main__":
    main()
