"""GPU parity of the one-kernel coupling network (csrc/coupling_nn.cu: conv -> ActNorm -> act -> conv1x1 -> ActNorm -> act ->
tap-split conv3x3, both hidden tensors in tensor memory, CTA pairs) against the oracle's conv chain and against the per-layer
kernels on the same inputs (Flow/glow_modules.py:229-240).

The kernel takes its weights with the ActNorm folded in (rows scaled by exp(logs) BEFORE the bf16 rounding, the shift as two
bf16 words inside the GEMM), so against the oracle and the per-layer kernels -- which round the weights first and scale the
fp32 accumulator afterwards -- the hidden activations differ by the bf16 rounding of the weights (2^-9 relative per weight,
averaging out over K) on top of fp32 accumulation order: 6e-3 of the output's max-norm."""
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


def make_case(B, Cin, C, H, W, hid, k, act, seed):
    g = torch.Generator().manual_seed(seed)
    x = bf(torch.randn(B, Cin, H, W, generator=g))
    w1 = bf(torch.randn(hid, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5)
    # ActNorm parameters (logs, bias): scale = exp(logs), shift = bias * scale  (Flow/glow_modules.py:40-45)
    l1, b1 = torch.randn(hid, generator=g) * 0.3, torch.randn(hid, generator=g) * 0.3
    w2 = bf(torch.randn(hid, hid, 1, 1, generator=g) / hid ** 0.5)
    l2, b2 = torch.randn(hid, generator=g) * 0.3, torch.randn(hid, generator=g) * 0.3
    w4 = bf(torch.randn(C, hid, 3, 3, generator=g) * 0.03)
    return x, w1, l1, b1, w2, l2, b2, w4


def folded(ops, w, logs, bias):
    hid = w.shape[0]
    return ops.pack_conv_weight_folded(w.cuda().contiguous(), logs.cuda().view(1, hid, 1, 1).contiguous(),
                                       bias.cuda().view(1, hid, 1, 1).contiguous())


def oracle_chain(x, w1, l1, b1, w2, l2, b2, w4, k, act):
    hid, C = w1.shape[0], w4.shape[0]
    s1, t1, s2, t2 = torch.exp(l1), b1 * torch.exp(l1), torch.exp(l2), b2 * torch.exp(l2)
    h1 = bf(O.act_fun(F.conv2d(x, w1, None, 1, (k - 1) // 2) * s1.view(1, hid, 1, 1) + t1.view(1, hid, 1, 1), act))
    h2 = bf(O.act_fun(F.conv2d(h1, w2) * s2.view(1, hid, 1, 1) + t2.view(1, hid, 1, 1), act))
    w9 = w4.permute(2, 3, 0, 1).reshape(9 * C, hid, 1, 1)
    return h1, h2, F.conv2d(h2, w9)


# B, Cin, C, H, W, hid, k
SHAPES = [
    (30, 18, 4, 32, 32, 256, 3),    # config J level 1 (W1 resident, 64-byte-row chunks), 240 tiles
    (9, 36, 8, 16, 16, 256, 3),     # level 2: W1 streams with the activations
    (5, 38, 12, 32, 32, 256, 3),    # config D level 1: 108 tap planes
    (7, 20, 6, 12, 20, 128, 3),     # ragged image, odd tile count, hidden 128
    (3, 256, 4, 32, 32, 256, 1),    # 1x1 first conv
    (1, 5, 2, 4, 4, 64, 3),         # a single, partly empty tile: the peer CTA's tile does not exist
    (9, 72, 16, 8, 8, 256, 3),      # config J level 3: 144 tap planes = two passes of GEMM3 (128 + 16 accumulator columns)
    (4, 76, 24, 16, 16, 256, 3),    # config D level 2: 216 tap planes
    (2, 40, 28, 8, 8, 192, 3),      # 252 tap planes, hidden 192
]


@pytest.mark.parametrize("act", ["relu", "leakyrelu"])
@pytest.mark.parametrize("B,Cin,C,H,W,hid,k", SHAPES)
def test_coupling_nn_fused_vs_oracle_and_layers(ops, act, B, Cin, C, H, W, hid, k):
    case = make_case(B, Cin, C, H, W, hid, k, act, seed=B * 100 + Cin)
    x, w1, l1, b1, w2, l2, b2, w4 = case
    h1_ref, h2_ref, taps_ref = oracle_chain(*case, k, act)
    w1f, cin_pad = folded(ops, w1, l1, b1)
    w2f, _ = folded(ops, w2, l2, b2)
    w9p, _ = ops.pack_tap_split_weight(w4.cuda())
    a = staged(ops, x)
    taps = torch.full((B, 9 * C, H, W), float("nan"), device="cuda")
    ops.coupling_nn_fused(a, cin_pad, k * k, w1f, hid, w2f, act, w9p, 9 * C, taps)
    assert max_rel(taps.cpu(), taps_ref) < 6e-3
    # per-layer kernels on the same inputs
    w1p, _ = ops.pack_conv_weight(w1.cuda())
    w2p, _ = ops.pack_conv_weight(w2.cuda())
    cu = [t.cuda() for t in (torch.exp(l1), b1 * torch.exp(l1), torch.exp(l2), b2 * torch.exp(l2))]
    hp = ops.cin_pad(hid)
    h1 = torch.zeros(B, H, W, hp, device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(a, cin_pad, w1p, hid, k * k, cu[0], cu[1], act, h1)
    h2 = torch.zeros(B, H, W, hp, device="cuda", dtype=torch.bfloat16)
    ops.conv_gemm(h1, hp, w2p, hid, 1, cu[2], cu[3], act, h2)
    taps2 = torch.empty(B, 9 * C, H, W, device="cuda")
    ops.conv_gemm(h2, hp, w9p, 9 * C, 1, None, None, "none", taps2)
    assert max_rel(taps.cpu(), taps2.cpu()) < 6e-3
    # side outputs (training): h1 / h2 leave by TMA store; the tap planes do not depend on the side outputs being requested
    taps3 = torch.full((B, 9 * C, H, W), float("nan"), device="cuda")
    h1s = torch.zeros(B, H, W, hp, device="cuda", dtype=torch.bfloat16)
    h2s = torch.zeros(B, H, W, hp, device="cuda", dtype=torch.bfloat16)
    ops.coupling_nn_fused(a, cin_pad, k * k, w1f, hid, w2f, act, w9p, 9 * C, taps3, h1s, h2s)
    assert torch.equal(taps3, taps)
    assert max_rel(h1s[..., :hid].permute(0, 3, 1, 2).float().cpu(), h1_ref) < 8e-3
    assert max_rel(h2s[..., :hid].permute(0, 3, 1, 2).float().cpu(), h2_ref) < 8e-3
    assert float(h1s[..., hid:].abs().max()) == 0 if hp > hid else True
    # the tap planes are exactly the tap-split conv of the h2 that was stored
    taps4 = torch.empty(B, 9 * C, H, W, device="cuda")
    ops.conv_gemm(h2s, hp, w9p, 9 * C, 1, None, None, "none", taps4)
    assert max_rel(taps.cpu(), taps4.cpu()) < 1e-4


def test_folded_weight_layout(ops):
    """rfk_pack_weight_folded: rows scaled by exp(logs), the shift as (hi, lo) bf16 words in the two columns behind K."""
    g = torch.Generator().manual_seed(3)
    w = torch.randn(64, 18, 3, 3, generator=g)
    logs, bias = torch.randn(64, generator=g) * 0.3, torch.randn(64, generator=g)
    wf, kp = folded(ops, w, logs, bias)
    assert kp == 32 and tuple(wf.shape) == (64, 9 * 32 + 16)
    main = wf[:, :288].float().cpu().view(64, 9, 32)
    ref = (w * torch.exp(logs).view(64, 1, 1, 1)).permute(0, 2, 3, 1).reshape(64, 9, 18)
    assert torch.equal(main[:, :, :18], ref.to(torch.bfloat16).float())
    assert float(main[:, :, 18:].abs().max()) == 0
    t = (bias * torch.exp(logs)).double()
    got = wf[:, 288].double().cpu() + wf[:, 289].double().cpu()
    assert float((got - t).abs().max() / t.abs().max()) < 2e-5
    assert float(wf[:, 290:].abs().max()) == 0


def test_coupling_nn_fused_many_tiles_per_pair(ops):
    """More tile pairs than CTA pairs (the role alternation of the two tensor-memory regions runs for many tiles), result
    checked on samples from both ends of the batch."""
    B, Cin, C, H, W, hid, k, act = 300, 18, 4, 32, 32, 256, 3, "relu"
    case = make_case(B, Cin, C, H, W, hid, k, act, seed=11)
    x, w1, l1, b1, w2, l2, b2, w4 = case
    w1f, cin_pad = folded(ops, w1, l1, b1)
    w2f, _ = folded(ops, w2, l2, b2)
    w9p, _ = ops.pack_tap_split_weight(w4.cuda())
    a = staged(ops, x)
    taps = torch.full((B, 9 * C, H, W), float("nan"), device="cuda")
    for _ in range(2):
        ops.coupling_nn_fused(a, cin_pad, 9, w1f, hid, w2f, act, w9p, 9 * C, taps)
    for sl in (slice(0, 2), slice(149, 151), slice(B - 2, B)):
        _, _, ref = oracle_chain(x[sl], w1, l1, b1, w2, l2, b2, w4, k, act)
        assert max_rel(taps[sl].cpu(), ref) < 6e-3


def test_scaled_dgrad_weight_layout(ops):
    """rfk_pack_weight_folded mode 5: the data-gradient weights with row r scaled by exp(logs[r]) equal the plain
    data-gradient packing of the weight whose input channel r was scaled before."""
    g = torch.Generator().manual_seed(4)
    w = torch.randn(256, 64, 3, 3, generator=g).cuda()
    logs = (torch.randn(1, 64, 1, 1, generator=g) * 0.3).cuda()
    got, kp = ops.pack_dgrad_weight_scaled(w, logs)
    ref, kp2 = ops.pack_dgrad_weight((w * torch.exp(logs)).contiguous())
    assert kp == kp2 and got.shape == ref.shape
    assert torch.equal(got, ref)
