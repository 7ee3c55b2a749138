"""GPU parity of the backward building blocks against torch CPU autograd on the oracle's formulation
(inputs and weights rounded to bf16 on both sides; tolerance 2e-3 of the reference's max-norm for fp32 outputs,
1e-2 where the kernel's output is bf16)."""
import pytest
import torch
import torch.nn.functional as F

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import recurrent_flows_msc_b200 as r
    return r.ops


def bf(x):
    return x.to(torch.bfloat16).float()


def max_rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def staged(ops, x):
    B, C, H, W = x.shape
    buf = torch.zeros(B, H, W, ops.cin_pad(C), device="cuda", dtype=torch.bfloat16)
    ops.pack_nhwc(x.cuda(), 0, C, buf, 0)
    return buf


CONVS = [(2, 18, 32, 32, 256, 3), (3, 256, 16, 16, 256, 1), (2, 64, 8, 8, 36, 1), (5, 72, 8, 8, 256, 3),
         (30, 288, 2, 2, 256, 3), (2, 40, 5, 7, 24, 3), (2, 256, 6, 6, 72, 1)]


@pytest.mark.parametrize("B,Cin,H,W,N,k", CONVS)
def test_conv_wgrad_and_dgrad(ops, B, Cin, H, W, N, k):
    g = torch.Generator().manual_seed(Cin + N)
    x = bf(torch.randn(B, Cin, H, W, generator=g)).requires_grad_(True)
    w = bf(torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).requires_grad_(True)
    dy = bf(torch.randn(B, N, H, W, generator=g))
    y = F.conv2d(x, w, None, 1, (k - 1) // 2)
    dx_ref, dw_ref = torch.autograd.grad(y, (x, w), dy)
    xs, dys = staged(ops, x.detach()), staged(ops, dy)
    dw = ops.conv_wgrad(xs, Cin, dys, N, k * k)
    assert dw.shape == w.shape
    assert max_rel(dw.cpu(), dw_ref) < 2e-3
    # data gradient = the forward conv kernel on tap-flipped, transposed weights
    wd, cpad = ops.pack_dgrad_weight(w.detach().cuda())
    dx = torch.empty(B, Cin, H, W, device="cuda")
    ops.conv_gemm(dys, cpad, wd, Cin, k * k, None, None, "none", dx)
    assert max_rel(dx.cpu(), dx_ref) < 2e-3


@pytest.mark.parametrize("act", ["relu", "leakyrelu", "none"])
@pytest.mark.parametrize("rows,n,ld", [(4096, 256, 256), (1000, 64, 64), (77, 16, 32), (3000, 512, 512)])
def test_act_affine_bwd(ops, act, rows, n, ld):
    g = torch.Generator().manual_seed(rows + n)
    a = torch.randn(rows, n, generator=g, requires_grad=True)
    logs = (torch.randn(n, generator=g) * 0.3).requires_grad_(True)
    bias = (torch.randn(n, generator=g) * 0.5).requires_grad_(True)
    v = (a + bias) * torch.exp(logs)                       # ActNorm on the conv output
    h = v if act == "none" else O.act_fun(v, act)
    h_b = bf(h.detach())
    dh = bf(torch.randn(rows, n, generator=g))
    # reference on the bf16-rounded output (what the kernel sees): recompute the mask / pre-activation from h_b
    slope = torch.ones_like(h_b) if act == "none" else torch.where(h_b > 0, 1.0, 0.2 if act == "leakyrelu" else 0.0)
    vv = h_b if act != "leakyrelu" else torch.where(h_b < 0, h_b * 5.0, h_b)
    dv = dh * slope
    s = torch.exp(logs.detach())
    da_ref, r_dv_ref, r_dvv_ref = dv * s, dv.sum(0), (dv * vv).sum(0)
    hb = torch.zeros(rows, ld, dtype=torch.bfloat16, device="cuda"); hb[:, :n] = h_b.to(torch.bfloat16).cuda()
    dhb = torch.zeros(rows, ld, dtype=torch.bfloat16, device="cuda"); dhb[:, :n] = dh.to(torch.bfloat16).cuda()
    da, r_dv, r_dvv = ops.act_affine_bwd(dhb, hb, n, s.cuda().contiguous(), act)
    assert max_rel(da[:, :n].float().cpu(), da_ref) < 1e-2
    assert max_rel(r_dv.cpu(), r_dv_ref) < 2e-3 and max_rel(r_dvv.cpu(), r_dvv_ref) < 2e-3
    # and those reductions ARE the ActNorm gradients (up to the bf16 rounding of h)
    dlogs_ref, dbias_ref = torch.autograd.grad(h, (logs, bias), dh)
    assert max_rel(r_dvv.cpu(), dlogs_ref) < 2e-2
    assert max_rel((s * r_dv.cpu()), dbias_ref) < 2e-2
