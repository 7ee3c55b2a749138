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
         (30, 288, 2, 2, 256, 3), (2, 40, 5, 7, 24, 3), (2, 256, 6, 6, 72, 1), (3, 7, 8, 8, 64, 3), (3, 11, 4, 4, 64, 3),
         (3, 64, 4, 4, 4, 3)]


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
@pytest.mark.parametrize("rows,n,ld", [(4096, 256, 256), (1000, 64, 64), (77, 16, 32), (3000, 512, 512), (300, 4, 32), (129, 13, 32)])
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


def gather_taps(taps, C):
    """S[c](y, x) = sum_t taps[t*C + c](y + ky - 1, x + kx - 1), zero outside the image."""
    H, W = taps.shape[2:]
    P = F.pad(taps, (1, 1, 1, 1))
    return sum(P[:, (3 * ky + kx) * C:(3 * ky + kx + 1) * C, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3))


@pytest.mark.parametrize("clamp", ["realnvp", "glow", "softclamp", "none"])
@pytest.mark.parametrize("B,C,H,W", [(3, 6, 16, 16), (2, 24, 8, 8), (5, 12, 5, 7), (2, 96, 2, 2)])
def test_coupling_taps_bwd(ops, clamp, B, C, H, W):
    g = torch.Generator().manual_seed(C * 7 + H)
    half = C // 2
    taps = (0.3 * torch.randn(B, 9 * C, H, W, generator=g)).requires_grad_()
    z = torch.randn(B, C, H, W, generator=g).requires_grad_()
    scale = (1 + 0.2 * torch.randn(C, generator=g)).requires_grad_()
    shift = (0.2 * torch.randn(C, generator=g)).requires_grad_()
    cs = (1 + 0.2 * torch.randn(half, generator=g)).requires_grad_()
    csh = (0.1 * torch.randn(half, generator=g)).requires_grad_()
    gz = torch.randn(B, C, H, W, generator=g)
    gld = torch.randn(B, generator=g)

    S = gather_taps(taps, C)
    S.retain_grad()
    h = S * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    t, raw = h[:, 0::2], h[:, 1::2]
    ls = O.clamp_log_scale(raw, clamp, cs, csh)
    z2o = (z[:, half:] + t) * torch.exp(ls)
    zo = torch.cat([z[:, :half], z2o], 1)
    ((zo * gz).sum() + (ls.sum(dim=(1, 2, 3)) * gld).sum()).backward()

    dz = gz.clone().cuda()
    dsum, d_scale, d_shift, d_cs, d_csh = ops.coupling_taps_bwd(
        taps.detach().cuda(), zo.detach().cuda().contiguous(), dz, scale.detach().cuda(), shift.detach().cuda(), clamp,
        cs.detach().cuda() if clamp == "realnvp" else None, csh.detach().cuda() if clamp == "realnvp" else None, gld.cuda())
    assert torch.equal(dz[:, :half].cpu(), gz[:, :half])          # z1 half untouched (its coupling-path gradient comes later)
    assert max_rel(dz[:, half:].cpu(), z.grad[:, half:]) < 2e-4
    assert max_rel(dsum.cpu(), S.grad) < 2e-4
    assert max_rel(d_scale.cpu(), scale.grad) < 2e-4
    assert max_rel(d_shift.cpu(), shift.grad) < 2e-4
    if clamp == "realnvp":
        assert max_rel(d_cs.cpu(), cs.grad) < 2e-4
        assert max_rel(d_csh.cpu(), csh.grad) < 2e-4
    # tap scatter = the gather's adjoint, as the NHWC bf16 operand of the tap GEMM's backward
    dt = ops.taps_scatter(dsum)
    assert dt.shape == (B, H, W, ops.pad_to(9 * C, 64))
    got = dt[..., :9 * C].float().permute(0, 3, 1, 2).cpu()
    assert max_rel(got, taps.grad) < 1e-2
    assert float(dt[..., 9 * C:].abs().max()) == 0.0 if dt.shape[-1] > 9 * C else True


@pytest.mark.parametrize("B,C,H,W", [(3, 12, 32, 32), (2, 6, 7, 5), (4, 48, 8, 8), (5, 64, 2, 2), (2, 24, 16, 16),
                                     (70, 12, 32, 32), (40, 4, 32, 32), (30, 8, 16, 16), (9, 32, 4, 4)])
def test_mix1x1_wgrad(ops, B, C, H, W):
    g = torch.Generator().manual_seed(C + B)
    x = torch.randn(B, C, H, W, generator=g)
    dy = torch.randn(B, C, H, W, generator=g)
    dW, db = ops.mix1x1_wgrad(x.cuda(), dy.cuda())
    ref_W = torch.einsum("bohw,bihw->oi", dy.double(), x.double())
    ref_b = dy.double().sum(dim=(0, 2, 3))
    assert max_rel(dW.cpu(), ref_W) < 1e-4
    assert max_rel(db.cpu(), ref_b) < 1e-4


@pytest.mark.parametrize("std_kind", ["exp", "softplus"])
@pytest.mark.parametrize("pairing", ["cross", "split"])
@pytest.mark.parametrize("with_params", [True, False])
def test_gauss_logp_bwd(ops, std_kind, pairing, with_params):
    import math
    import recurrent_flows_msc_b200 as r
    g = torch.Generator().manual_seed(11)
    B, zC, off, n, H, W = 3, 10, 4, 6, 5, 7
    z = torch.randn(B, zC, H, W, generator=g).requires_grad_()
    params = (0.5 * torch.randn(B, 2 * n, H, W, generator=g)).requires_grad_() if with_params else None
    gb = torch.randn(B, generator=g)
    if with_params:
        mean, raw = (params[:, 0::2], params[:, 1::2]) if pairing == "cross" else (params[:, :n], params[:, n:])
    else:
        mean, raw = torch.zeros(B, n, H, W), torch.zeros(B, n, H, W)
    sd = torch.exp(raw) if std_kind == "exp" else F.softplus(raw) + 1e-8
    zz = z[:, off:off + n]
    lp = (-0.5 * math.log(2 * math.pi) - torch.log(sd) - 0.5 * ((zz - mean) / sd) ** 2).sum(dim=(1, 2, 3))
    (lp * gb).sum().backward()
    dz0 = torch.randn(B, zC, H, W, generator=g)
    dz = dz0.clone().cuda()
    pair = r._lib.PAIR_CROSS if pairing == "cross" else r._lib.PAIR_SPLIT
    dp = ops.gauss_logp_bwd(z.detach().cuda(), off, n, params.detach().cuda() if with_params else None, pair, std_kind,
                            gb.cuda(), dz)
    assert max_rel(dz.cpu() - dz0, z.grad) < 1e-4
    if with_params:
        assert max_rel(dp.cpu(), params.grad) < 1e-4
    else:
        assert dp is None


@pytest.mark.parametrize("B,Cin,H,W,N,with_perm", [(2, 18, 32, 32, 256, True), (3, 7, 8, 8, 64, False), (2, 36, 16, 16, 256, True),
                                                   (2, 40, 5, 7, 24, False)])
def test_dgrad_tap_split(ops, B, Cin, H, W, N, with_perm):
    """Data gradient of a 3x3 conv as ONE 1x1 GEMM with N = 9*Cin (bf16 NHWC planes) + the shifted nine-plane gather."""
    g = torch.Generator().manual_seed(Cin * 3 + N)
    w = bf(0.1 * torch.randn(N, Cin, 3, 3, generator=g))
    dy = bf(torch.randn(B, N, H, W, generator=g))
    x = torch.zeros(B, Cin, H, W, requires_grad=True)
    F.conv2d(x, w, padding=1).backward(dy)
    perm = torch.randperm(Cin, generator=g) if with_perm else None    # staging channel j = weight channel perm[j]
    ref = x.grad if perm is None else x.grad[:, perm]
    wd9, cp, r8 = ops.pack_dgrad_taps_weight(w.cuda(), None if perm is None else perm.cuda())
    dys = staged(ops, dy)
    planes = torch.empty(B, H, W, ops.pad_to(9 * r8, 64), device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(dys, cp, wd9, 9 * r8, 1, None, None, "none", planes)
    out = torch.empty(B, Cin, H, W, device="cuda")
    ops.taps_gather_nhwc(planes, Cin, r8, out)
    assert max_rel(out.cpu(), ref) < 1e-2     # the nine planes are rounded to bf16
    # accumulate form: channels [0, n0) += into one tensor, the rest += into the first channels of another
    for n0 in (0, Cin // 3, Cin):
        acc0 = torch.randn(B, n0 + 2, H, W, generator=g).cuda()
        acc1 = torch.randn(B, Cin - n0 + 3, H, W, generator=g).cuda()
        a0, a1 = acc0.clone(), acc1.clone()
        ops.taps_gather_nhwc_acc(planes, Cin, r8, acc0, n0, acc1)
        assert torch.equal(acc0[:, :n0], a0[:, :n0] + out[:, :n0]) and torch.equal(acc0[:, n0:], a0[:, n0:])
        assert torch.equal(acc1[:, :Cin - n0], a1[:, :Cin - n0] + out[:, n0:]) and torch.equal(acc1[:, Cin - n0:], a1[:, Cin - n0:])
