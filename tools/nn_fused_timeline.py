"""Cycle accounting of the one-kernel coupling network (debug aid, needs a GPU): python tools/nn_fused_timeline.py [store]
MMA thread (leader CTAs): cycles per tile pair in GEMM1 issue + waits, waiting for TMA data, for the drained tap accumulator,
for the first / later chunks of h1 and h2.  Epilogue warp 0: waiting for the three accumulators and busy in each epilogue."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from recurrent_flows_msc_b200 import ops, _lib

store = len(sys.argv) > 1 and sys.argv[1] == "store"
B, hw, cin, C = (int(v) for v in os.environ.get("SHAPE", "570,32,18,4").split(","))
x = torch.randn(B, cin, hw, hw, device="cuda")
act = torch.zeros(B, hw, hw, ops.cin_pad(cin), device="cuda", dtype=torch.bfloat16)
ops.pack_nhwc(x, 0, cin, act, 0)
an = lambda: (torch.randn(1, 256, 1, 1, device="cuda") * 0.1, torch.randn(1, 256, 1, 1, device="cuda") * 0.1)
w1f, cp = ops.pack_conv_weight_folded(torch.randn(256, cin, 3, 3, device="cuda") / (cin * 9) ** 0.5, *an())
w2f, _ = ops.pack_conv_weight_folded(torch.randn(256, 256, 1, 1, device="cuda") / 16, *an())
w9p, _ = ops.pack_tap_split_weight(torch.randn(C, 256, 3, 3, device="cuda") * 0.03)
h1 = torch.zeros(B, hw, hw, 256, device="cuda", dtype=torch.bfloat16)
h2 = torch.zeros_like(h1)
taps = torch.zeros(B, 9 * C, hw, hw, device="cuda")
run = lambda: ops.coupling_nn_fused(act, cp, 9, w1f, 256, w2f, "relu", w9p, 9 * C, taps, *((h1, h2) if store else ()))
for _ in range(3):
    run()
ncta = 148
tl = torch.zeros(2 * ncta * 16, dtype=torch.int64, device="cuda")
_lib.call("rfk_debug_set_timeline", tl.data_ptr(), 2 * ncta)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
_lib.call("rfk_debug_set_timeline", None, 0)
stamps = tl.view(2 * ncta, 16)[ncta:].cpu().double()[0::2]
full = tl.view(2 * ncta, 16)[:ncta].cpu().double()
lead = full[0::2]
tiles = torch.full((ncta // 2,), float(-(-(-(-B * hw * hw // 128) + 1) // 2 // (ncta // 2))), dtype=torch.float64)
print(f"kernel {e0.elapsed_time(e1) * 1e3:.1f} us (instrumented), store={store}, {tiles.median():.0f} tile pairs per CTA pair")
names = ["mma: GEMM1 section (issue + data waits)", "mma:   of which waiting for TMA data", "mma: wait tap accumulator drained",
         "mma: wait h1 chunk 0", "mma: wait h1 chunks 1-3", "mma: wait h2 chunk 0", "mma: wait h2 chunks 1-3", "mma: whole loop",
         "epi: wait GEMM1 accumulator", "epi: epilogue 1 busy", "epi: wait GEMM2 accumulator", "epi: epilogue 2 busy",
         "epi: wait GEMM3 accumulator", "epi: epilogue 3 busy", "epi:   of which tcgen05.ld + wait", "epi: whole loop"]
for i, nme in enumerate(names):
    print(f"  cycles / tile pair  {nme:42s} leader {(lead[:, i] / tiles).median():8.0f}   peer {(full[1::2][:, i] / tiles).median():8.0f}")

labels = ["mma: GEMM1 issue starts", "mma: GEMM1 issued", "mma: h1 chunk 0 seen", "mma: h1 chunk 3 seen", "mma: GEMM2 issued",
          "mma: h2 chunk 0 seen", "mma: h2 chunk 3 seen", "mma: GEMM3 issued", "epi: GEMM1 accumulator seen", "epi: epilogue 1 done",
          "epi: GEMM2 accumulator seen", "epi: epilogue 2 done", "epi: GEMM3 accumulator seen", "epi: epilogue 3 done"]
rel = (stamps[:, :14] - stamps[:, :1]) % (2 ** 32)
order = sorted(range(14), key=lambda i: float(rel[:, i].median()))
print("tile 5 of every leader CTA, SM cycles since its GEMM1 issue started (median over CTAs):")
for i in order:
    print(f"  {rel[:, i].median():8.0f}  {labels[i]}")
