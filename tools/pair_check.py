"""Correctness + timing of the CTA-pair conv kernel vs the single-CTA build (RFK_GEMM_PAIR read per process: run twice)."""
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

torch.manual_seed(0)
for (B, hw, cin, taps) in [(570, 32, 18, 9), (570, 32, 256, 1), (570, 32, 4, 9), (569, 32, 18, 9), (570, 16, 36, 9)]:
    k = 3 if taps == 9 else 1
    x = torch.randn(B, cin, hw, hw, device="cuda")
    act = torch.zeros(B, hw, hw, ops.cin_pad(cin), device="cuda", dtype=torch.bfloat16)
    ops.pack_nhwc(x, 0, cin, act, 0)
    w = torch.randn(256, cin, k, k, device="cuda") / (cin * taps) ** 0.5
    wp, cp = ops.pack_conv_weight(w)
    sc, sh = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda") * 0.1
    out = torch.zeros(B, hw, hw, 256, device="cuda", dtype=torch.bfloat16)
    us = t(lambda: ops.conv_gemm(act, cp, wp, 256, taps, sc, sh, "relu", out))
    xb, wb = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
    sub = slice(0, 8)
    ref = torch.relu(torch.nn.functional.conv2d(xb[sub], wb, padding=k // 2) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    got = out[sub].float().permute(0, 3, 1, 2)
    err = float((got - ref).abs().max() / ref.abs().max())
    ref2 = torch.relu(torch.nn.functional.conv2d(xb[-4:], wb, padding=k // 2) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    err2 = float((out[-4:].float().permute(0, 3, 1, 2) - ref2).abs().max() / ref2.abs().max())
    print(f"PAIR={os.environ.get('RFK_GEMM_PAIR', '1')} B={B} {hw}x{hw} cin={cin} taps={taps}: {us:7.1f} us  rel err first/last samples {err:.2e} {err2:.2e}")
