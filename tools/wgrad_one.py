"""One rfk_conv_wgrad shape for ncu: python tools/wgrad_one.py hw cin cout taps"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recurrent_flows_msc_b200 import ops
hw, cin, cout, taps = (int(a) for a in sys.argv[1:5])
B = 570
x = torch.randn(B, hw, hw, ops.cin_pad(cin), device="cuda").to(torch.bfloat16)
dy = torch.randn(B, hw, hw, ops.cin_pad(cout), device="cuda").to(torch.bfloat16)
k = 3 if taps == 9 else 1
dw = torch.zeros(cout, cin, k, k, device="cuda")
for _ in range(2):
    ops.conv_wgrad(x, cin, dy, cout, taps, out=dw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.conv_wgrad(x, cin, dy, cout, taps, out=dw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
