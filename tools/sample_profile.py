"""One RFN.predict inner step (ConvLSTM cell + ListGlow.sample, 30 sequences, config J) bracketed by cudaProfilerStart/Stop
for `ncu --profile-from-start off`; prints the eager and CUDA-graph step times."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import recurrent_flows_msc_b200 as rf

B = 30
torch.manual_seed(0)
flow = rf.ListGlow([B, 1, 64, 64], bench.cond_sizes(bench.J, B), [B, 256, 2, 2], bench.glow_args(bench.J)).eval()
bench.trained_like(flow, 0)
flow = flow.cuda()
_, conds, _, _ = bench.synth_inputs(bench.J, B, 1, 1)
base = torch.randn(B, bench.J["base_ch"], 2, 2)
conds = [c.cuda() for c in conds]
base = base.cuda()
with torch.no_grad():
    for _ in range(3):
        flow.sample(None, conds, base, num_samples=B, temperature=0.7)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    flow.sample(None, conds, base, num_samples=B, temperature=0.7)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    g = rf.GraphedSample(flow, conds, base, temperature=0.7)
    for _ in range(3):
        g(conds, base)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        g(conds, base)
    b.record()
    torch.cuda.synchronize()
    print(f"graph replay {a.elapsed_time(b) / 20:.3f} ms per step of {B} frames")
