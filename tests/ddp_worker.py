"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU, NCCL).

Checks, on a small ListGlow + ConvLSTM training step:
  1. the all-reduced flat gradient (sum over ranks) / world == the single-GPU gradient of the concatenated batch;
  2. replicas hold bit-identical parameters after 5 steps of FlatAdam (eager and CUDA-graph replay)."""
import math
import os
import sys
import types

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import recurrent_flows_msc_b200 as rf
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    a = types.SimpleNamespace(LU_decomposed=True, n_units_affine=64, non_lin_glow="relu", clamp_type="realnvp",
                              flow_norm="actnorm", flow_batchnorm_momentum=0.0, learn_prior=True, n_units_prior=32,
                              make_conditional=True, base_norm="actnorm", split2d_act="softplus", L=2, K=2, n_bits=8)
    Bs = 4                                   # per-rank batch
    Bg = Bs * world
    cond = lambda n: [[n, 8, 8, 8], [n, 16, 4, 4]]   # noqa: E731

    def build(n, ws):
        torch.manual_seed(0)
        m = rf.ListGlow([n, 1, 16, 16], cond(n), [n, 12, 4, 4], a).train()
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for name, p in m.named_parameters():
                p.add_(torch.randn(p.shape, generator=g) * (0.03 if "conv.weight" in name else 0.1))
            for name, b in m.named_buffers():
                if name.endswith("initialized"):
                    b.fill_(1)
        lstm = rf.ConvLSTM(6, 8, [3, 3]).train()
        m, lstm = m.to(dev), lstm.to(dev)
        opt = rf.FlatAdam(list(m.parameters()) + list(lstm.parameters()), lr=1e-3, world_size=ws)
        ranges = opt.attach(m)      # per-level slices all-reduced during the backward sweep
        assert set(ranges) == {0, 1, "prior"}, ranges
        return m, lstm, opt

    g = torch.Generator().manual_seed(7)
    x = (torch.floor(torch.rand(Bg, 1, 16, 16, generator=g) * 256) / 256 - 0.5).to(dev)
    noise = (torch.rand(Bg, 1, 16, 16, generator=g) / 256).to(dev)
    conds = [torch.randn(*s, generator=g).to(dev) for s in cond(Bg)]
    feats = torch.randn(Bg, 3, 6, 4, 4, generator=g).to(dev)
    zpart = torch.randn(Bg, 4, 4, 4, generator=g).to(dev)

    def loss_of(m, lstm, sl):
        hs, h, _ = lstm(feats[sl])
        base = torch.cat([h, zpart[sl]], 1)
        _, nll = m.log_prob(x[sl], [c[sl] for c in conds], base, noise=noise[sl])
        return nll.mean() / (math.log(2.0) * 256)

    def mark(msg):
        print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)

    mark("data ready")
    # ---- 1. gradient equality --------------------------------------------------------------------------------------
    m, lstm, opt = build(Bs, world)
    sl = slice(rank * Bs, (rank + 1) * Bs)
    opt.zero_grad()
    loss_of(m, lstm, sl).backward()
    opt.gather_grads()
    opt.allreduce_grads()
    g_dp = opt.flat_g.clone() / world
    mark("all-reduced gradient ready")
    m1, lstm1, opt1 = build(Bg, 1)
    opt1.zero_grad()
    loss_of(m1, lstm1, slice(0, Bg)).backward()
    opt1.gather_grads()
    g_single = opt1.flat_g
    err = float((g_dp - g_single).abs().max() / g_single.abs().max())
    cos = float(torch.dot(g_dp.double(), g_single.double()) / (g_dp.double().norm() * g_single.double().norm()))
    # ---- 2. replicas identical after 5 steps (eager, then graph replay) ---------------------------------------------
    for _ in range(3):
        opt.zero_grad()
        loss_of(m, lstm, sl).backward()
        opt.step()
    mark("eager steps done")
    step = rf.GraphedTrainStep(lambda: loss_of(m, lstm, sl), opt, warmup=1)
    mark("captured: " + step.mode)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    mark("graph replays done")
    mine = opt.flat_p.clone()
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    same = all(torch.equal(allp[0], q) for q in allp)
    moved = float((mine - opt1.flat_p).abs().max())
    ok = err < 2e-3 and cos > 0.9999 and same and moved > 0 and bool(torch.isfinite(mine).all())
    if rank == 0:
        print(f"ddp_worker: world {world}: allreduced-gradient vs single-GPU max-norm rel err {err:.3e}, cosine {cos:.6f}; "
              f"replicas bit-identical after 5 steps: {same}; parameters moved {moved:.3e}; {'OK' if ok else 'FAIL'}", flush=True)
    # the captured graph holds NCCL work: release it before the communicator goes away (destroying the process group with
    # a live graph that contains a collective hung at exit on the 2-GPU box)
    del step
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    mark("exiting")
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
