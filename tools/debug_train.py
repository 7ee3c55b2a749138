"""Debug aid: checks every conv weight gradient of the training path against torch on the tape's own operands."""
import math, sys, types
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import recurrent_flows_msc_b200 as rf
from recurrent_flows_msc_b200 import ops
from recurrent_flows_msc_b200.Flow import training as T

orig = T._conv_bwd


def checked(st, mod, x_act, cin, da, perm=None, dgrad_out=None, key="id"):
    n = mod.conv.out_channels
    before = st.grads.get(id(mod.conv.weight))
    before = None if before is None else before.clone()
    orig(st, mod, x_act, cin, da, perm, dgrad_out, key)
    got = st.grads[id(mod.conv.weight)] - (0 if before is None else before)
    x = x_act[..., :cin].float().permute(0, 3, 1, 2).contiguous()
    if perm is not None:
        inv = torch.empty_like(perm); inv[perm] = torch.arange(perm.numel(), device=perm.device)
        x = x[:, inv]
    dy = da[..., :n].float().permute(0, 3, 1, 2).contiguous()
    w = mod.conv.weight.detach().float().requires_grad_()
    with torch.enable_grad():
        y = F.conv2d(x, w, padding=mod.conv.padding)
        (gw,) = torch.autograd.grad(y, w, dy)
    err = float((got - gw).abs().max() / gw.abs().max().clamp_min(1e-30))
    print(f"wgrad cin={cin} n={n} taps={mod.taps} perm={perm is not None} rel_err={err:.2e} max={float(gw.abs().max()):.3e}")


T._conv_bwd = checked
ARGS = dict(LU_decomposed=True, n_units_affine=64, non_lin_glow="relu", clamp_type="realnvp", flow_norm="actnorm",
            flow_batchnorm_momentum=0.0, learn_prior=True, n_units_prior=32, make_conditional=True, base_norm="actnorm",
            split2d_act="softplus", L=2, K=2, n_bits=8)
B = 3
torch.manual_seed(1)
m = rf.ListGlow([B, 1, 16, 16], [[B, 5, 8, 8], [B, 7, 4, 4]], [B, 6, 4, 4], types.SimpleNamespace(**ARGS)).train()
gen = torch.Generator().manual_seed(1)
with torch.no_grad():
    for name, p in m.named_parameters():
        p.add_(torch.randn(p.shape, generator=gen) * (0.03 if "conv.weight" in name else 0.1))
    for name, b in m.named_buffers():
        if name.endswith("initialized"):
            b.fill_(1)
m = m.cuda()
x = (torch.rand(B, 1, 16, 16) - 0.5).cuda()
conds = [torch.randn(B, 5, 8, 8).cuda(), torch.randn(B, 7, 4, 4).cuda()]
base = torch.randn(B, 6, 4, 4).cuda()
z, nll = m.log_prob(x, conds, base, logdet=0)
nll.mean().backward()
