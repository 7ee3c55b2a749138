"""Achieved GB/s of the bandwidth-bound kernels on >L2 tensors (same function bench.py reports as roofline.elementwise_at_scale)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
peak = 6547.8
try:
    peak = json.load(open(os.path.join(bench.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except (OSError, KeyError):
    pass
for e in bench.elementwise_at_scale(torch.device("cuda", 0), peak):
    print(f"{e['kernel']:40s} {e['algorithmic_mb']:8.1f} MB {e['us']:8.1f} us {e['gbs']:8.1f} GB/s  {100 * e['hbm_frac']:5.1f}% of {peak:.0f}")
