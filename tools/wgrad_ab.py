import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recurrent_flows_msc_b200 import ops
B = 570
for (hw, cin, cout) in [(32, 18, 256), (32, 256, 4), (32, 256, 18), (32, 4, 256), (16, 256, 8)]:
    x = torch.randn(B, hw, hw, ops.cin_pad(cin), device="cuda").to(torch.bfloat16)
    dy = torch.randn(B, hw, hw, ops.cin_pad(cout), device="cuda").to(torch.bfloat16)
    dw = torch.zeros(cout, cin, 3, 3, device="cuda")
    for _ in range(3):
        ops.conv_wgrad(x, cin, dy, cout, 9, out=dw)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10):
        ops.conv_wgrad(x, cin, dy, cout, 9, out=dw)
    b.record(); torch.cuda.synchronize()
    print(f"quads_off={os.environ.get('RFK_WGRAD_NO_QUADS','0')} hw={hw} cin={cin} cout={cout}: {a.elapsed_time(b)*100:.1f} us")
