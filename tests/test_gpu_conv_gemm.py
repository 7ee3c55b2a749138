"""GPU parity of the tcgen05 implicit-GEMM convolution against the CPU oracle's conv (through the C ABI).

Inputs and weights are rounded to bf16 on both sides, so the only difference left is fp32
accumulation order: the tolerance is 2e-3 of the output's max-norm (bf16-output cases add the
bf16 rounding of the result, 2^-8 relative)."""
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


CASES = [
    # B, Cin, H, W, N, k
    (2, 64, 8, 16, 16, 1),
    (2, 9, 6, 6, 16, 3),
    (3, 18, 32, 32, 256, 3),
    (5, 36, 16, 16, 256, 3),
    (3, 256, 32, 32, 256, 1),
    (30, 288, 2, 2, 256, 3),
    (7, 144, 4, 4, 256, 3),
    (3, 72, 8, 8, 40, 3),
    (2, 130, 5, 7, 24, 3),
    (1, 64, 64, 64, 64, 3),
    (2, 256, 2, 2, 512, 3),
    (2, 100, 12, 20, 384, 1),
]


@pytest.mark.parametrize("B,Cin,H,W,N,k", CASES)
def test_conv_gemm_f32_out(ops, B, Cin, H, W, N, k):
    g = torch.Generator().manual_seed(B * 1000 + Cin)
    x = bf(torch.randn(B, Cin, H, W, generator=g))
    w = bf(torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    scale, shift = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
    ref = F.conv2d(x, w, None, 1, (k - 1) // 2) * scale.view(1, N, 1, 1) + shift.view(1, N, 1, 1)
    wp, cin_pad = ops.pack_conv_weight(w.cuda())
    out = torch.full((B, N, H, W), float("nan"), device="cuda")
    ops.conv_gemm(staged(ops, x), cin_pad, wp, N, k * k, scale.cuda(), shift.cuda(), "none", out)
    assert max_rel(out.cpu(), ref) < 2e-3
    raw = torch.empty(B, N, H, W, device="cuda")
    ops.conv_gemm(staged(ops, x), cin_pad, wp, N, k * k, None, None, "relu", raw)
    assert max_rel(raw.cpu(), F.relu(F.conv2d(x, w, None, 1, (k - 1) // 2))) < 2e-3


@pytest.mark.parametrize("B,Cin,H,W,N,k", CASES[:8])
@pytest.mark.parametrize("act", ["relu", "leakyrelu"])
def test_conv_gemm_bf16_nhwc_out(ops, B, Cin, H, W, N, k, act):
    g = torch.Generator().manual_seed(7 + N)
    x = bf(torch.randn(B, Cin, H, W, generator=g))
    w = bf(torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    scale, shift = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.2
    ref = O.act_fun(F.conv2d(x, w, None, 1, (k - 1) // 2) * scale.view(1, N, 1, 1) + shift.view(1, N, 1, 1), act)
    wp, cin_pad = ops.pack_conv_weight(w.cuda())
    ld = ops.pad_to(N + 8, 64)
    out = torch.zeros(B, H, W, ld, device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(staged(ops, x), cin_pad, wp, N, k * k, scale.cuda(), shift.cuda(), act, out, out_off=8)
    got = out[..., 8:8 + N].permute(0, 3, 1, 2).float().cpu()
    assert max_rel(got, ref) < 6e-3
    assert float(out[..., :8].abs().max()) == 0 and float(out[..., 8 + N:].abs().max()) == 0


@pytest.mark.parametrize("clamp", ["realnvp", "glow", "softclamp", "none"])
@pytest.mark.parametrize("B,C,H,W", [(3, 4, 32, 32), (2, 8, 16, 16), (5, 16, 8, 8), (9, 32, 4, 4), (33, 64, 2, 2), (2, 12, 6, 10)])
def test_conv_gemm_coupling(ops, clamp, B, C, H, W):
    g = torch.Generator().manual_seed(C)
    hid = 64
    h2 = bf(torch.relu(torch.randn(B, hid, H, W, generator=g)))
    w = bf(torch.randn(C, hid, 3, 3, generator=g) * 0.02)
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    cs, csh = torch.randn(C // 2, generator=g) * 0.5, torch.randn(C // 2, generator=g) * 0.1
    z = torch.randn(B, C, H, W, generator=g)
    nn_out = F.conv2d(h2, w, None, 1, 1) * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1)
    sh, raw = nn_out[:, 0::2], nn_out[:, 1::2]
    ls = O.clamp_log_scale(raw, clamp, cs, csh)
    z_ref = torch.cat([z[:, :C // 2], (z[:, C // 2:] + sh) * torch.exp(ls)], 1)
    wp, cin_pad = ops.pack_conv_weight(w.cuda())
    zc, ld = z.cuda().clone(), torch.zeros(B, device="cuda")
    act = staged(ops, h2)
    ops.conv_gemm_coupling(act, cin_pad, wp, C, 9, scale.cuda(), shift.cuda(), zc, clamp, cs.cuda(), csh.cuda(), ld, False)
    assert max_rel(zc.cpu(), z_ref) < 2e-3
    torch.testing.assert_close(ld.cpu(), ls.sum(dim=(1, 2, 3)), rtol=2e-3, atol=2e-2)
    ops.conv_gemm_coupling(act, cin_pad, wp, C, 9, scale.cuda(), shift.cuda(), zc, clamp, cs.cuda(), csh.cuda(), ld, True)
    assert max_rel(zc.cpu(), z) < 1e-4          # exact inverse with the same NN output
    assert float(ld.abs().max()) < 1e-3


@pytest.mark.parametrize("B,Cin,Hc,H,W", [(2, 3, 4, 5, 6), (2, 64, 64, 16, 16), (3, 20, 60, 8, 8), (4, 40, 200, 2, 2), (30, 512, 200, 2, 2), (5, 300, 24, 3, 3)])
def test_conv_gemm_lstm(ops, B, Cin, Hc, H, W):
    import recurrent_flows_msc_b200 as r
    g = torch.Generator().manual_seed(Hc)
    cell = r.ConvLSTMLayer(Cin, Hc, [3, 3], True).cuda()
    with torch.no_grad():
        cell.conv[0].weight.copy_(bf(cell.conv[0].weight))
    x, h0, c0 = bf(torch.randn(B, Cin, H, W, generator=g)), bf(torch.randn(B, Hc, H, W, generator=g)), torch.randn(B, Hc, H, W, generator=g)
    w, b = cell.conv[0].weight.detach().cpu(), cell.conv[0].bias.detach().cpu()
    h_ref, c_ref = O.convlstm_cell(x, h0, c0, w, b)
    with torch.no_grad():
        h, c = cell(x.cuda(), [h0.cuda(), c0.cuda()])
    assert max_rel(c.cpu(), c_ref) < 2e-3
    assert max_rel(h.cpu(), h_ref) < 2e-3
    h_ref0, c_ref0 = O.convlstm_cell(x, None, None, w, b)
    with torch.no_grad():
        h1, c1 = cell(x.cuda(), [None, None])
    assert max_rel(c1.cpu(), c_ref0) < 2e-3 and max_rel(h1.cpu(), h_ref0) < 2e-3


@pytest.mark.parametrize("clamp", ["realnvp", "glow"])
@pytest.mark.parametrize("B,C,H,W", [(3, 4, 32, 32), (2, 8, 16, 16), (5, 16, 8, 8), (2, 12, 6, 10), (3, 28, 3, 5)])
def test_tap_split_coupling(ops, clamp, B, C, H, W):
    """The tap-split form (1x1 GEMM with N=9C + shifted-plane gather) equals the 3x3 conv + coupling tail."""
    g = torch.Generator().manual_seed(C + 100)
    hid = 64
    h2 = bf(torch.relu(torch.randn(B, hid, H, W, generator=g)))
    w = bf(torch.randn(C, hid, 3, 3, generator=g) * 0.02)
    scale, shift = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    cs, csh = torch.randn(C // 2, generator=g) * 0.5, torch.randn(C // 2, generator=g) * 0.1
    z = torch.randn(B, C, H, W, generator=g)
    nn_out = F.conv2d(h2, w, None, 1, 1) * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1)
    ls = O.clamp_log_scale(nn_out[:, 1::2], clamp, cs, csh)
    z_ref = torch.cat([z[:, :C // 2], (z[:, C // 2:] + nn_out[:, 0::2]) * torch.exp(ls)], 1)
    w9, cin_pad = ops.pack_tap_split_weight(w.cuda())
    taps = torch.empty(B, 9 * C, H, W, device="cuda")
    ops.conv_gemm(staged(ops, h2), cin_pad, w9, 9 * C, 1, None, None, "none", taps)
    zc, ld = z.cuda().clone(), torch.zeros(B, device="cuda")
    ops.coupling_tail_taps(taps, zc, scale.cuda(), shift.cuda(), clamp, cs.cuda(), csh.cuda(), ld, False)
    assert max_rel(zc.cpu(), z_ref) < 2e-3
    torch.testing.assert_close(ld.cpu(), ls.sum(dim=(1, 2, 3)), rtol=2e-3, atol=2e-2)
    ops.coupling_tail_taps(taps, zc, scale.cuda(), shift.cuda(), clamp, cs.cuda(), csh.cuda(), ld, True)
    assert max_rel(zc.cpu(), z) < 1e-4
    assert float(ld.abs().max()) < 1e-3


@pytest.mark.parametrize("B,Cin,H,W,N,ks", [(30, 288, 2, 2, 256, 9), (7, 144, 4, 4, 256, 9), (40, 128, 4, 4, 256, 3),
                                            (3, 512, 2, 2, 512, 9), (5, 130, 3, 5, 40, 3)])
def test_conv_gemm_splitk_fused(ops, B, Cin, H, W, N, ks):
    """Split-K with the in-kernel last-arriver fix-up equals the single-pass kernel; run twice to prove that workspace
    and counters are left clean."""
    g = torch.Generator().manual_seed(Cin + N)
    x = bf(torch.randn(B, Cin, H, W, generator=g))
    w = bf(torch.randn(N, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5)
    scale, shift = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.2
    ref = torch.relu(F.conv2d(x, w, None, 1, 1) * scale.view(1, N, 1, 1) + shift.view(1, N, 1, 1))
    wp, cin_pad = ops.pack_conv_weight(w.cuda())
    ld = ops.pad_to(N, 64)
    for _ in range(2):
        out = torch.zeros(B, H, W, ld, device="cuda", dtype=torch.bfloat16)
        ops.conv_gemm_splitk_fused(staged(ops, x), cin_pad, wp, N, 9, ks, scale.cuda(), shift.cuda(), "relu", out)
        got = out[..., :N].permute(0, 3, 1, 2).float().cpu()
        assert max_rel(got, ref) < 6e-3
        assert float(out[..., N:].abs().max()) == 0 if ld > N else True


@pytest.mark.parametrize("act", ["relu", "leakyrelu"])
@pytest.mark.parametrize("B,C,H,W,hid", [(3, 4, 32, 32, 256), (2, 8, 16, 16, 256), (5, 14, 8, 8, 128), (2, 6, 6, 10, 64),
                                         (40, 4, 2, 2, 192)])
def test_conv1x1_taps_fused(ops, act, B, C, H, W, hid):
    """conv1x1 -> affine -> act -> tap-split conv in one kernel (h2 in tensor memory) equals the two-kernel path and
    the oracle's conv chain."""
    g = torch.Generator().manual_seed(C * 7 + hid)
    h1 = bf(torch.relu(torch.randn(B, hid, H, W, generator=g)))
    w2 = bf(torch.randn(hid, hid, 1, 1, generator=g) / hid ** 0.5)
    s2, t2 = torch.rand(hid, generator=g) + 0.5, torch.randn(hid, generator=g) * 0.2
    w4 = bf(torch.randn(C, hid, 3, 3, generator=g) * 0.03)
    h2_ref = bf(O.act_fun(F.conv2d(h1, w2) * s2.view(1, hid, 1, 1) + t2.view(1, hid, 1, 1), act))
    w9 = w4.permute(2, 3, 0, 1).reshape(9 * C, hid, 1, 1)
    taps_ref = F.conv2d(h2_ref, w9)
    w2p, cin_pad = ops.pack_conv_weight(w2.cuda())
    w9p, _ = ops.pack_tap_split_weight(w4.cuda())
    a = staged(ops, h1)
    taps = torch.full((B, 9 * C, H, W), float("nan"), device="cuda")
    ops.conv1x1_taps_fused(a, cin_pad, w2p, hid, s2.cuda(), t2.cuda(), act, w9p, 9 * C, taps)
    assert max_rel(taps.cpu(), taps_ref) < 4e-3
    # two-kernel path on the same inputs
    h2 = torch.zeros(B, H, W, ops.cin_pad(hid), device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(a, cin_pad, w2p, hid, 1, s2.cuda(), t2.cuda(), act, h2)
    taps2 = torch.empty(B, 9 * C, H, W, device="cuda")
    ops.conv_gemm(h2, ops.cin_pad(hid), w9p, 9 * C, 1, None, None, "none", taps2)
    assert max_rel(taps.cpu(), taps2.cpu()) < 1e-3


@pytest.mark.parametrize("B,Cin,H,W,N,k,act", [(40, 18, 32, 32, 256, 3, "relu"), (38, 256, 32, 32, 256, 1, "leakyrelu"),
                                               (593, 20, 8, 8, 256, 3, "none"), (75, 4, 32, 32, 224, 3, "relu")])
def test_conv_gemm_cta_pair_shapes(ops, B, Cin, H, W, N, k, act):
    """Shapes with at least two pixel tiles per SM and resident weights: these launch as clusters of two CTAs
    (tcgen05 cta_group::2, weights split over the pair); 593 x 8x8 gives an ODD tile count (the odd CTA's last tile does
    not exist).  The reference is evaluated on a few samples at both ends of the batch."""
    g = torch.Generator().manual_seed(B + Cin)
    x = bf(torch.randn(B, Cin, H, W, generator=g))
    w = bf(torch.randn(N, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    scale, shift = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.3
    wp, cin_pad = ops.pack_conv_weight(w.cuda())
    ld = ops.pad_to(N, 64)
    out = torch.zeros(B, H, W, ld, device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(staged(ops, x), cin_pad, wp, N, k * k, scale.cuda(), shift.cuda(), act, out)
    for sl in (slice(0, 3), slice(B // 2, B // 2 + 2), slice(B - 3, B)):
        ref = F.conv2d(x[sl], w, None, 1, (k - 1) // 2) * scale.view(1, N, 1, 1) + shift.view(1, N, 1, 1)
        ref = ref if act == "none" else O.act_fun(ref, act)
        got = out[sl][..., :N].permute(0, 3, 1, 2).float().cpu()
        assert max_rel(got, ref) < 6e-3
    assert float(out[..., N:].abs().max()) == 0 if ld > N else True
