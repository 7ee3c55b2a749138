"""Timing of the one-kernel coupling network (rfk_coupling_nn_fused) against the per-layer launches at the large levels of
config J (570 frames): python tools/nn_fused_bench.py [one]   ("one": a single fused launch at level 1, for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recurrent_flows_msc_b200 import ops


def t(fn, n=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / n


def setup(B, hw, cin, C, hid=256):
    x = torch.randn(B, cin, hw, hw, device="cuda")
    act = torch.zeros(B, hw, hw, ops.cin_pad(cin), device="cuda", dtype=torch.bfloat16)
    ops.pack_nhwc(x, 0, cin, act, 0)
    w1 = torch.randn(hid, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5
    w2 = torch.randn(hid, hid, 1, 1, device="cuda") / hid ** 0.5
    w1p, cp = ops.pack_conv_weight(w1)
    w2p, _ = ops.pack_conv_weight(w2)
    an = lambda: (torch.randn(1, hid, 1, 1, device="cuda") * 0.1, torch.randn(1, hid, 1, 1, device="cuda") * 0.1)
    w1f, w2f = ops.pack_conv_weight_folded(w1, *an())[0], ops.pack_conv_weight_folded(w2, *an())[0]
    w9p, _ = ops.pack_tap_split_weight(torch.randn(C, hid, 3, 3, device="cuda") * 0.03)
    aff = [torch.rand(hid, device="cuda") + 0.5, torch.randn(hid, device="cuda") * 0.1,
           torch.rand(hid, device="cuda") + 0.5, torch.randn(hid, device="cuda") * 0.1]
    h1 = torch.zeros(B, hw, hw, hid, device="cuda", dtype=torch.bfloat16)
    h2 = torch.zeros(B, hw, hw, hid, device="cuda", dtype=torch.bfloat16)
    taps = torch.zeros(B, 9 * C, hw, hw, device="cuda")
    return act, cp, w1p, w2p, w9p, aff, h1, h2, taps, w1f, w2f


torch.manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] == "one":
    store = len(sys.argv) > 2 and sys.argv[2] == "store"
    act, cp, w1p, w2p, w9p, aff, h1, h2, taps, w1f, w2f = setup(570, 32, 18, 4)
    run = lambda: ops.coupling_nn_fused(act, cp, 9, w1f, 256, w2f, "relu", w9p, 36, taps, *((h1, h2) if store else ()))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)

for (B, hw, cin, C) in [(570, 32, 18, 4), (570, 16, 36, 8), (30, 32, 18, 4), (32, 32, 38, 12)]:
    act, cp, w1p, w2p, w9p, aff, h1, h2, taps, w1f, w2f = setup(B, hw, cin, C)
    n3 = 9 * C
    M = B * hw * hw
    gf = 2.0 * M * 256 * (9 * cin + 256 + n3) / 1e6   # MFLOP, so that MFLOP / us = TFLOP/s

    def layers_infer():
        ops.conv_gemm(act, cp, w1p, 256, 9, aff[0], aff[1], "relu", h1)
        ops.conv1x1_taps_fused(h1, 256, w2p, 256, aff[2], aff[3], "relu", w9p, n3, taps)

    def layers_train():
        ops.conv_gemm(act, cp, w1p, 256, 9, aff[0], aff[1], "relu", h1)
        ops.conv_gemm(h1, 256, w2p, 256, 1, aff[2], aff[3], "relu", h2)
        ops.conv_gemm(h2, 256, w9p, n3, 1, None, None, "none", taps)

    fused = lambda: ops.coupling_nn_fused(act, cp, 9, w1f, 256, w2f, "relu", w9p, n3, taps)
    fused_s = lambda: ops.coupling_nn_fused(act, cp, 9, w1f, 256, w2f, "relu", w9p, n3, taps, h1, h2)
    a, b, c, d = t(layers_infer), t(layers_train), t(fused), t(fused_s)
    print(f"B={B} {hw}x{hw} cin={cin} C={C} ({gf / 1e3:.1f} GFLOP): per-layer inference {a:7.1f} us, per-layer training {b:7.1f} us, "
          f"fused {c:7.1f} us ({gf / c:.0f} TFLOP/s), fused + h1/h2 stores {d:7.1f} us", flush=True)
