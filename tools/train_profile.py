"""Per-kernel time of one ListGlow training step (log_prob forward with tape + hand-written backward) at bench.py's shape.
Usage: python tools/train_profile.py [n_frames]   (default 570)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import recurrent_flows_msc_b200 as rf  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 570
dev = torch.device("cuda", 0)
torch.manual_seed(0)
flow = rf.ListGlow([n, 1, 64, 64], bench.cond_sizes(n), [n, bench.J["base_ch"], 2, 2], bench.glow_args()).train()
bench.trained_like(flow, 0)
flow = flow.to(dev)
x, conds, base, _ = bench.synth_inputs(n, 1, 1)
x, base, conds = x.to(dev), base.to(dev), [c.to(dev) for c in conds]
opt = rf.FlatAdam(flow.parameters(), lr=1e-4)


def loss_fn():
    _, nll = flow.log_prob(x, conds, base)
    return nll.mean() / (0.6931 * 4096)


def step():
    opt.zero_grad()
    loss = loss_fn()
    loss.backward()
    return loss


def timed(fn, n=3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, float(out.detach())   # no reference to the autograd graph survives


for _ in range(2):
    step()
torch.cuda.synchronize()
print("peak memory GB (eager)", torch.cuda.max_memory_allocated() / 2**30)
ms, loss = timed(step)
print(f"eager fwd+bwd {ms:.2f} ms   loss {loss:.4f}")


def full():
    loss = step()
    opt.step()
    return loss


if os.environ.get("NCU") == "1":   # ncu --profile-from-start off: exactly one eager training step (incl. repack + Adam)
    full()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    full()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
ms, loss = timed(full)
print(f"eager fwd+bwd+FlatAdam (weights repacked every step) {ms:.2f} ms  loss {loss:.4f}")
print("peak memory GB", torch.cuda.max_memory_allocated() / 2**30)
if os.environ.get("GRAPH", "1") == "1":
    g = rf.GraphedTrainStep(loss_fn, opt, warmup=2)
    ms, loss = timed(g, 5)
    print(f"graphed train step {ms:.2f} ms -> {n / ms * 1e3:.0f} frames/s   loss {loss:.4f}")
    print("peak memory GB", torch.cuda.max_memory_allocated() / 2**30)
t0 = rf._lib.launches
kt = bench.KernelTimer()
rf._lib.tracer = kt
step()
rf._lib.tracer = None
agg = kt.table()
print("launches", rf._lib.launches - t0)
tot = sum(d["ms"] for d in agg.values())
for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    tf = d["flops"] / d["ms"] / 1e9 if d["flops"] else 0
    print(f"{k:28s} {d['launches']:5d} {d['ms']:8.3f} ms {100 * d['ms'] / tot:5.1f}%  {tf:7.1f} TF/s")
print("sum of kernel times", tot)
for k, d in sorted(kt.shapes.items(), key=lambda kv: -kv[1]["ms"])[:14]:
    print(k, d["launches"], f"{d['ms']:.3f} ms", f"{d['flops'] / d['ms'] / 1e9:.1f} TF/s")
