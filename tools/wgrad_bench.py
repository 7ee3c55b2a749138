"""Times rfk_conv_wgrad on the training step's shapes (config J, 570 frames).  env RFK_WGRAD_ATOMIC_BUDGET etc. apply."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recurrent_flows_msc_b200 as rf
from recurrent_flows_msc_b200 import ops

B = 570
shapes = []
for l, (C, cc) in enumerate(zip([4, 8, 16, 32, 64], [16, 32, 64, 128, 256])):
    hw = 32 >> l
    cin = C // 2 + cc
    shapes += [(hw, cin, 256, 9, "net0"), (hw, 256, 256, 1, "net2"), (hw, 256, C, 9, "net4")]
tot = 0.0
for hw, cin, cout, taps, name in shapes:
    x = torch.randn(B, hw, hw, ops.cin_pad(cin), device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, hw, hw, ops.cin_pad(cout), device="cuda").to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    dw = torch.zeros(cout, cin, k, k, device="cuda")
    for _ in range(3):
        ops.conv_wgrad(x, cin, dy, cout, taps, out=dw)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        ops.conv_wgrad(x, cin, dy, cout, taps, out=dw)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 100
    tot += us
    fl = 2.0 * B * hw * hw * cin * cout * taps
    print(f"{name} hw={hw:2d} cin={cin:3d} cout={cout:3d} taps={taps}  {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s")
print(f"total per GlowStep-set {tot:.1f} us  (x10 steps per level = {tot / 100:.2f} ms per training step)")
