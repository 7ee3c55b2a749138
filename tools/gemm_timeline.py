"""Per-CTA phase timeline of the conv-GEMM kernel (debug aid, needs a GPU).
usage: python tools/gemm_timeline.py [taps cin n]   (default: the level-1 1x1 conv of config J)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import recurrent_flows_msc_b200 as rf
from recurrent_flows_msc_b200 import ops, _lib

taps, cin, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1, 256, 256)
B, H, W = (int(v) for v in os.environ.get('BHW', '570,32,32').split(','))
k = 3 if taps == 9 else 1
act = torch.randn(B, H, W, ops.cin_pad(cin), device="cuda").to(torch.bfloat16)
w = torch.randn(n, cin, k, k, device="cuda") * 0.05
wp, cin_pad = ops.pack_conv_weight(w)
out = torch.zeros(B, H, W, ops.pad_to(n, 64), device="cuda", dtype=torch.bfloat16)
scale, shift = torch.ones(n, device="cuda"), torch.zeros(n, device="cuda")
KS = int(os.environ.get("SPLITK", "0"))
def run():
    if KS:
        ops.conv_gemm_splitk_fused(act, cin_pad, wp, n, taps, KS, scale, shift, "relu", out)
    else:
        ops.conv_gemm(act, cin_pad, wp, n, taps, scale, shift, "relu", out)
for _ in range(3):
    run()
ncta = 148 * 4
tl = torch.zeros(ncta * 16, dtype=torch.int64, device="cuda")
_lib.call("rfk_debug_set_timeline", tl.data_ptr(), ncta)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
_lib.call("rfk_debug_set_timeline", None, 0)
full = tl.view(ncta, 16).cpu().double()
full = full[full[:, 7] > 0]
if os.environ.get('EVEN') == '1':
    full = full[0::2]   # CTA-pair mode: the leader CTAs hold the MMA-side counters
ncta = full.shape[0]
t = full[:, :8]
t0 = t[:, 0].min()
names = ["start", "setup", "wres", "tmaN", "mmaN", "acc0", "epi0", "done"]
print(f"kernel {e0.elapsed_time(e1)*1e3:.1f} us, {ncta} CTAs; span {(t.max()-t0)/1e3:.1f} us")
d = t - t[:, :1]
for i, nme in enumerate(names):
    print(f"  {nme:8s} since CTA start: median {d[:, i].median()/1e3:8.2f} us  p10 {d[:, i].quantile(0.1)/1e3:8.2f}  p90 {d[:, i].quantile(0.9)/1e3:8.2f}")
tiles = max(1.0, (B * H * W / 128) / ncta)
print(f"{ncta} CTAs ran")
for i, nme in [(8, "producer wait free stage"), (9, "mma wait data"), (10, "mma wait drained accumulator"),
               (11, "epilogue wait accumulator"), (12, "epilogue busy"), (13, "mma issue section"),
               (14, "epi: tmem ld+wait"), (15, "epi: math+release"), (2, "epi: wait prev store + bar"), (5, "epi: sts+fence+bar"), (6, "epi: tma store issue")]:
    print(f"  cycles/tile {nme:30s} {full[:, i].median() / tiles:9.0f}")
print(f"per-tile steady state ~ {(d[:, 7].median() - d[:, 6].median()) / 1e3 / (tiles - 1):.2f} us ({tiles:.1f} tiles per CTA)")
