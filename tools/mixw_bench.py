import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recurrent_flows_msc_b200 import ops
for (C, hw) in [(4, 32), (8, 16), (16, 8), (32, 4), (64, 2), (12, 32), (48, 8)]:
    x = torch.randn(570, C, hw, hw, device="cuda"); dy = torch.randn_like(x)
    for _ in range(3): ops.mix1x1_wgrad(x, dy)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): ops.mix1x1_wgrad(x, dy)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    print(f"blocked={os.environ.get('RFK_MIXW_BLOCKED','1')} C={C} {hw}x{hw}: {a.elapsed_time(b)*50:.1f} us (incl. the zero-fill of the result)")
