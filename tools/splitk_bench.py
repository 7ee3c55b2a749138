"""Times conv_gemm vs conv_gemm_splitk_fused on small-M 3x3 layers (sampling shapes)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recurrent_flows_msc_b200 import ops

def t(fn, n=50):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / n

for (B, hw, cin) in [(30, 2, 288), (30, 4, 144), (30, 8, 72), (570, 2, 288)]:
    act = torch.randn(B, hw, hw, ops.cin_pad(cin), device="cuda").to(torch.bfloat16)
    w = torch.randn(256, cin, 3, 3, device="cuda") * 0.05
    wp, cp = ops.pack_conv_weight(w)
    out = torch.zeros(B, hw, hw, 256, device="cuda", dtype=torch.bfloat16)
    sc, sh = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
    base = t(lambda: ops.conv_gemm(act, cp, wp, 256, 9, sc, sh, "relu", out))
    res = [f"plain {base:6.1f} us"]
    for ks in (3, 9):
        res.append(f"k_split={ks} {t(lambda: ops.conv_gemm_splitk_fused(act, cp, wp, 256, 9, ks, sc, sh, 'relu', out)):6.1f} us")
    print(f"B={B} {hw}x{hw} cin={cin}: " + "  ".join(res))
