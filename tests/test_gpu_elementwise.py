"""GPU parity of the bandwidth-bound kernels against the CPU oracle (through the C ABI)."""
import pytest
import torch

import oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import recurrent_flows_msc_b200 as r
    return r.ops


def rel_err(a, b):
    return float((a.cpu().double() - b.cpu().double()).abs().max() / b.cpu().double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(2, 3, 4, 6), (3, 1, 64, 64), (2, 4, 32, 32), (5, 16, 2, 2), (1, 2, 8, 24), (2, 3, 6, 10)])
def test_squeeze_bit_exact(ops, shape):
    x = torch.randn(*shape)
    y = ops.squeeze2d(x.cuda(), False)
    assert torch.equal(y.cpu(), O.squeeze2d(x, False))
    assert torch.equal(ops.squeeze2d(y, True).cpu(), x)


def test_squeeze_golden(ops):
    g = load_golden("squeeze")
    assert torch.equal(ops.squeeze2d(g["x"].cuda(), False).cpu(), g["y"])
    assert torch.equal(ops.squeeze2d(g["y"].cuda(), True).cpu(), g["undo"])


@pytest.mark.parametrize("shape", [(3, 5, 4, 4), (2, 7, 3, 5), (4, 256, 32, 32)])
def test_actnorm_and_init(ops, shape):
    x = torch.randn(*shape) * 1.7 + 0.3
    C = shape[1]
    bias = torch.zeros(1, C, 1, 1, device="cuda")
    logs = torch.zeros(1, C, 1, 1, device="cuda")
    ops.actnorm_init(x.cuda(), bias, logs)
    b_ref, l_ref = O.actnorm_init(x)
    torch.testing.assert_close(bias.cpu(), b_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(logs.cpu(), l_ref, rtol=1e-5, atol=1e-6)
    y = ops.actnorm(x.cuda(), bias, logs, False)
    y_ref, _ = O.actnorm(x, b_ref, l_ref, None, False)
    torch.testing.assert_close(y.cpu(), y_ref, rtol=1e-5, atol=1e-5)
    xr = ops.actnorm(y, bias, logs, True)
    torch.testing.assert_close(xr.cpu(), x, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,C,H,W", [(2, 4, 32, 32), (3, 6, 3, 5), (2, 64, 2, 2), (2, 12, 8, 8), (1, 192, 2, 2)])
def test_mix1x1(ops, B, C, H, W):
    x, Wm, bv = torch.randn(B, C, H, W), torch.randn(C, C) / C ** 0.5, torch.randn(C)
    ref = torch.einsum("oi,bihw->bohw", Wm.double(), x.double()) + bv.double().view(1, C, 1, 1)
    side = torch.zeros(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    y = ops.mix1x1(x.cuda(), Wm.cuda(), bv.cuda(), side=side, side_n=C // 2 if C // 2 <= 60 else 0, side_off=3)
    assert rel_err(y, ref) < 1e-5
    if C // 2 <= 60:
        want = ref[:, :C // 2].permute(0, 2, 3, 1).float().to(torch.bfloat16)
        assert torch.equal(side[..., 3:3 + C // 2].cpu(), y[:, :C // 2].permute(0, 2, 3, 1).to(torch.bfloat16).cpu())
        assert rel_err(side[..., 3:3 + C // 2].float(), want.float()) < 1e-2
        assert float(side[..., :3].abs().max()) == 0 and float(side[..., 3 + C // 2:].abs().max()) == 0
    y2 = ops.mix1x1(x.cuda(), Wm.cuda())
    assert rel_err(y2, ref - bv.double().view(1, C, 1, 1)) < 1e-5


def test_pack_and_copy(ops):
    x = torch.randn(3, 10, 5, 6)
    dst = torch.zeros(3, 5, 6, 64, device="cuda", dtype=torch.bfloat16)
    ops.pack_nhwc(x.cuda(), 2, 7, dst, 8)
    want = x[:, 2:9].permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(dst[..., 8:15].cpu(), want)
    assert float(dst[..., :8].abs().max()) == 0 and float(dst[..., 15:].abs().max()) == 0
    ops.pack_nhwc(x.cuda(), 0, 10, dst, 17)  # unaligned offset -> scalar path
    assert torch.equal(dst[..., 17:27].cpu(), x.permute(0, 2, 3, 1).to(torch.bfloat16))
    seq = torch.randn(2, 4, 3, 4, 4).cuda()  # batch-strided slice
    d2 = torch.zeros(2, 4, 4, 64, device="cuda", dtype=torch.bfloat16)
    ops.pack_nhwc(seq[:, 2], 0, 3, d2, 0)
    assert torch.equal(d2[..., :3].cpu(), seq[:, 2].permute(0, 2, 3, 1).to(torch.bfloat16).cpu())
    out = torch.zeros(3, 12, 5, 6, device="cuda")
    ops.copy_channels(x.cuda(), 1, out, 4, 6)
    assert torch.equal(out[:, 4:10].cpu(), x[:, 1:7]) and float(out[:, :4].abs().max()) == 0


@pytest.mark.parametrize("clamp", ["realnvp", "glow", "softclamp", "none"])
@pytest.mark.parametrize("shape", [(2, 8, 6, 6), (3, 4, 32, 32), (2, 6, 3, 5)])
def test_coupling_tail(ops, clamp, shape):
    B, C, H, W = shape
    nn_out, z = torch.randn(*shape), torch.randn(*shape)
    cs, csh = torch.randn(C // 2) * 0.5, torch.randn(C // 2) * 0.1
    shift, raw = nn_out[:, 0::2], nn_out[:, 1::2]
    ls = O.clamp_log_scale(raw, clamp, cs, csh)
    z_ref = torch.cat([z[:, :C // 2], (z[:, C // 2:] + shift) * torch.exp(ls)], 1)
    ld_ref = ls.sum(dim=(1, 2, 3))
    zc, ld = z.cuda().clone(), torch.zeros(B, device="cuda")
    ops.coupling_tail(nn_out.cuda(), zc, clamp, cs.cuda(), csh.cuda(), ld, False)
    torch.testing.assert_close(zc.cpu(), z_ref, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ld.cpu(), ld_ref, rtol=1e-4, atol=1e-3)
    ops.coupling_tail(nn_out.cuda(), zc, clamp, cs.cuda(), csh.cuda(), ld, True)
    torch.testing.assert_close(zc.cpu(), z, rtol=1e-4, atol=1e-4)
    assert float(ld.abs().max()) < 1e-3


@pytest.mark.parametrize("pairing,std_kind", [(0, "softplus"), (0, "exp"), (1, "exp")])
def test_gauss(ops, pairing, std_kind):
    B, n, H, W = 3, 5, 4, 6
    z, params, eps = torch.randn(B, 2 * n, H, W), torch.randn(B, 2 * n, H, W), torch.randn(B, n, H, W)
    mean, raw = (params[:, 0::2], params[:, 1::2]) if pairing == 0 else (params[:, :n], params[:, n:])
    std = torch.nn.functional.softplus(raw) + 1e-8 if std_kind == "softplus" else torch.exp(raw)
    ref = torch.distributions.Normal(mean, std).log_prob(z[:, n:]).sum(dim=(1, 2, 3))
    ld = torch.ones(B, device="cuda")
    ops.gauss_logp(z.cuda(), n, params.cuda(), n, pairing, std_kind, ld)
    torch.testing.assert_close(ld.cpu(), ref + 1, rtol=1e-4, atol=1e-3)
    ld0 = torch.zeros(B, device="cuda")
    ops.gauss_logp(z.cuda(), 0, None, n, pairing, "exp", ld0)  # null params = N(0,1)
    torch.testing.assert_close(ld0.cpu(), torch.distributions.Normal(0., 1.).log_prob(z[:, :n]).sum(dim=(1, 2, 3)),
                               rtol=1e-4, atol=1e-3)
    out = torch.zeros(B, 2 * n, H, W, device="cuda")
    ops.gauss_sample(eps.cuda(), params.cuda(), n, pairing, std_kind, 0.7, out, n)
    torch.testing.assert_close(out[:, n:].cpu(), mean + std * 0.7 * eps, rtol=1e-5, atol=1e-5)
    assert float(out[:, :n].abs().max()) == 0


@pytest.mark.parametrize("shape", [(2, 4, 5, 6), (2, 64, 16, 16), (3, 5, 3, 3)])
def test_lstm_pointwise(ops, shape):
    B, Hc, H, W = shape
    cc, c = torch.randn(B, 4 * Hc, H, W), torch.randn(B, Hc, H, W)
    peep = torch.randn(3, Hc, H, W) * 0.3
    i = torch.sigmoid(cc[:, :Hc] + peep[0] * c)
    f = torch.sigmoid(cc[:, Hc:2 * Hc] + peep[1] * c)
    g = torch.tanh(cc[:, 3 * Hc:])
    cn = f * c + i * g
    o = torch.sigmoid(cc[:, 2 * Hc:3 * Hc] + peep[2] * cn)
    h, c2 = ops.convlstm_pointwise(cc.cuda(), c.cuda(), peep.cuda())
    torch.testing.assert_close(c2.cpu(), cn, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(h.cpu(), o * torch.tanh(cn), rtol=1e-5, atol=1e-5)
    h0, c0 = ops.convlstm_pointwise(cc.cuda(), c.cuda(), None)
    i0, f0 = torch.sigmoid(cc[:, :Hc]), torch.sigmoid(cc[:, Hc:2 * Hc])
    cn0 = f0 * c + i0 * g
    torch.testing.assert_close(c0.cpu(), cn0, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(h0.cpu(), torch.sigmoid(cc[:, 2 * Hc:3 * Hc]) * torch.tanh(cn0), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("B,C,H,W", [(3, 4, 32, 32), (2, 8, 16, 16), (5, 16, 8, 8), (7, 32, 4, 4), (9, 64, 2, 2), (2, 6, 3, 5)])
def test_coupling_taps_mix_equals_gather_then_mix(ops, reverse, B, C, H, W):
    """The fused tap-gather + coupling tail + 1x1 mix equals the two separate kernels (which are checked against the
    oracle elsewhere), including both log-det contributions and the bf16 side output."""
    g = torch.Generator().manual_seed(C + 31 * int(reverse))
    taps = (torch.randn(B, 9 * C, H, W, generator=g) * 0.2).cuda()
    z = torch.randn(B, C, H, W, generator=g).cuda()
    scale, shift = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    cs, csh = (torch.randn(C // 2, generator=g) * 0.5).cuda(), (torch.randn(C // 2, generator=g) * 0.1).cuda()
    Wm, bv = (torch.randn(C, C, generator=g) / C ** 0.5).cuda(), torch.randn(C, generator=g).cuda()
    addend = torch.tensor([0.37], device="cuda")
    # reference composition
    z_ref, ld_ref = z.clone(), torch.zeros(B, device="cuda")
    ops.coupling_tail_taps(taps, z_ref, scale, shift, "realnvp", cs, csh, ld_ref, reverse)
    side_ref = torch.zeros(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    y_ref = ops.mix1x1(z_ref, Wm, bv, side=side_ref, side_n=C // 2 if C <= 64 else 0, side_off=8, logdet=ld_ref,
                       addend=addend, alpha=-1.0 if reverse else 1.0)
    # fused
    ld = torch.zeros(B, device="cuda")
    side = torch.zeros(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    z_in = z.clone()
    y = ops.coupling_taps_mix(taps, z_in, scale, shift, "realnvp", cs, csh, ld, reverse, Wm, bv, side=side,
                              side_n=C // 2 if C <= 64 else 0, side_off=8, logdet=ld, addend=addend,
                              alpha=-1.0 if reverse else 1.0)
    assert torch.equal(z_in, z), "input must not be modified"
    torch.testing.assert_close(y, y_ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(ld, ld_ref, rtol=1e-4, atol=1e-3)
    assert rel_err(side.float(), side_ref.float()) < 1e-2
