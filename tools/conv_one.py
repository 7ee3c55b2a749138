"""One rfk_conv_gemm launch (bf16 NHWC out, ReLU) for ncu: python tools/conv_one.py B hw cin taps"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recurrent_flows_msc_b200 import ops
B, hw, cin, taps = (int(a) for a in sys.argv[1:5])
k = 3 if taps == 9 else 1
act = torch.randn(B, hw, hw, ops.cin_pad(cin), device="cuda").to(torch.bfloat16)
w = torch.randn(256, cin, k, k, device="cuda") * 0.05
wp, cp = ops.pack_conv_weight(w)
out = torch.zeros(B, hw, hw, 256, device="cuda", dtype=torch.bfloat16)
sc, sh = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
for _ in range(2):
    ops.conv_gemm(act, cp, wp, 256, taps, sc, sh, "relu", out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.conv_gemm(act, cp, wp, 256, taps, sc, sh, "relu", out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
