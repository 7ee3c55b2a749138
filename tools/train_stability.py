"""Runs N graphed training steps of the bench's hot path and prints the loss trajectory (finite, decreasing)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import recurrent_flows_msc_b200 as rf
n, steps = 570, int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
torch.manual_seed(0)
flow = rf.ListGlow([n, 1, 64, 64], bench.cond_sizes(bench.J, n), [n, bench.J["base_ch"], 2, 2], bench.glow_args(bench.J)).train()
bench.trained_like(flow, 0)
flow = flow.to(dev)
x, conds, _, _ = bench.synth_inputs(bench.J, n, 1, 1)
base = torch.randn(n, bench.J["base_ch"], 2, 2)
x, base, conds = x.to(dev), base.to(dev), [c.to(dev) for c in conds]
opt = rf.FlatAdam(flow.parameters(), lr=1e-4)
def loss_fn():
    _, nll = flow.log_prob(x, conds, base)
    return nll.mean() / (math.log(2.0) * 4096)
step = rf.GraphedTrainStep(loss_fn, opt, warmup=2)
losses = [float(step()) for _ in range(steps)]
print("bits/dim every 10 steps:", [round(v, 2) for v in losses[::10]], "last", round(losses[-1], 3))
assert all(math.isfinite(v) for v in losses), "non-finite loss"
assert losses[-1] < losses[0], "loss did not decrease"
with torch.no_grad():
    _, nll = flow.eval().log_prob(x, conds, base)
print("eval-mode bits/dim after training:", round(float(nll.mean() / (math.log(2.0) * 4096)), 3))
