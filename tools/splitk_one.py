import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recurrent_flows_msc_b200 import ops
B, hw, cin = 30, 2, 288
act = torch.randn(B, hw, hw, ops.cin_pad(cin), device="cuda").to(torch.bfloat16)
w = torch.randn(256, cin, 3, 3, device="cuda") * 0.05
wp, cp = ops.pack_conv_weight(w)
out = torch.zeros(B, hw, hw, 256, device="cuda", dtype=torch.bfloat16)
sc, sh = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
for _ in range(3):
    ops.conv_gemm(act, cp, wp, 256, 9, sc, sh, "relu", out)
    ops.conv_gemm_splitk_fused(act, cp, wp, 256, 9, 9, sc, sh, "relu", out)
    ops.conv_gemm_splitk_fused(act, cp, wp, 256, 9, 3, sc, sh, "relu", out)
torch.cuda.synchronize()
