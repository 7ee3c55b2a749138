"""Diagnose stock-vs-patched ListGlow.sample differences inside the reference's RFN (run on the GPU box)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ref_helpers import job_script_args, reference_dir, stub_optional_imports  # noqa: E402

REF = reference_dir()
stub_optional_imports()
sys.path.insert(0, REF)
import Flow  # noqa: E402
import Utils  # noqa: E402
rfn_mod = importlib.import_module("RFN.RFN_new")
import recurrent_flows_msc_b200 as rfk  # noqa: E402

B, T = 4, 20
args = job_script_args(REF, B)
torch.manual_seed(0)
stock = rfn_mod.RFN(args).cuda().train()
scale = float(os.environ.get("PERT", "1.0"))
with torch.no_grad():
    g = torch.Generator().manual_seed(5)
    for n, p in stock.named_parameters():
        if n.startswith("flow."):
            p.add_((torch.randn(p.shape, generator=g) * scale * (0.01 if "conv.weight" in n else 0.05)).cuda())
g = torch.Generator().manual_seed(0)
u = torch.rand(B, T, 1, 64, 64, generator=g) * (torch.rand(B, T, 1, 64, 64, generator=g) < 0.3).float()
x = (torch.floor(u * 256) / 256 - 0.5).cuda()
torch.manual_seed(11)
stock.loss(x, 0)
sd1 = {k: v.clone() for k, v in stock.state_dict().items()}
rfk.install_into(Flow, Utils)
rfn_mod2 = importlib.reload(rfn_mod)
torch.manual_seed(0)
ours = rfn_mod2.RFN(args).cuda()
ours.load_state_dict(sd1)
stock.eval(); ours.eval()


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


with torch.no_grad():
    gg = torch.Generator().manual_seed(3)
    conds = [torch.randn(B, c, 32 >> l, 32 >> l, generator=gg).cuda() for l, c in enumerate([16, 32, 64, 128, 256])]
    base = torch.randn(B, 256, 2, 2, generator=gg).cuda()
    for temp in (1e-6, 0.7):
        torch.manual_seed(21)
        xs = stock.flow.sample(None, conds, base, temperature=temp)
        torch.manual_seed(21)
        xo = ours.flow.sample(None, conds, base, temperature=temp)
        torch.manual_seed(22)
        xs2 = stock.flow.sample(None, conds, base, temperature=temp)
        print(f"T={temp}: stock vs ours {rel(xo, xs):.3e}; stock vs stock(other seed) {rel(xs2, xs):.3e}; |x| max {float(xs.abs().max()):.3f}")
    # forward then reverse on the same z: f/g consistency of each implementation and cross check
    xin = x[:, 3]
    zs, nll_s = stock.flow.log_prob(xin, conds, base, 0)
    zo, nll_o = ours.flow.log_prob(xin, conds, base, 0)
    print("log_prob z err", rel(zo, zs), "nll", float(nll_s.mean()), float(nll_o.mean()))
    hs = stock.lstm(torch.randn(B, 1, 512, 2, 2).cuda().mul(0).add(1.0), stock.h_0, stock.c_0)[1]
    ho = ours.lstm(torch.ones(B, 1, 512, 2, 2).cuda(), ours.h_0, ours.c_0)[1]
    print("lstm h err", rel(ho, hs))
    torch.manual_seed(21)
    ts, ps = stock.predict(x, 3, 10)
    torch.manual_seed(21)
    to, po = ours.predict(x, 3, 10)
    for i in range(3):
        print("predict frame", i, rel(po[i], ps[i]))
