#!/usr/bin/env python
"""Benchmark of the RFN hot path (Glow decoder + ConvLSTM recurrence) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload = "rfn_J_hotpath_fwd"): the hot path of ONE RFN training batch of the
reference's job-script configuration J (RFN/default_rfn_job.sh: B=30 sequences, 1x64x64, 10+10 frames,
L=5, K=10, hidden 256, h=200, z=56):
  * 19 ConvLSTM cell steps (512 -> 200 hidden channels, 3x3, 2x2 maps, batch 30), and
  * ListGlow.log_prob (dequantise, f: 5 levels x 10 GlowSteps + 4 Split2d, learned prior) on the
    B*(T-1) = 570 predicted frames, time-batched into one call (exact because nothing the
    flow produces feeds back into the recurrence).
One step = one such batch, forward direction (density evaluation).  Backward kernels do not exist yet, so
this is NOT a full training step; the JSON line says so in config.pass.  value = frames / s with inputs
resident in HBM; e2e = same through the public nn.Module API from pinned HOST buffers (H2D of x, the
condition pyramid and the ConvLSTM input, D2H of nll and h) inside the timed region.

With --gpus N (torchrun) every rank processes its own batch of 30 sequences (weak scaling, no data-path
collective); time = max over ranks.
"""
import argparse
import contextlib
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "RFN 64x64 frames/sec (Glow decoder + ConvLSTM hot path, forward)"
J = dict(B=30, T=20, L=5, K=10, hidden=256, n_units_prior=512, cond_ch=[16, 32, 64, 128, 256],
         base_ch=256, lstm_in=512, lstm_hidden=200, n_bits=8)


def glow_args():
    return types.SimpleNamespace(LU_decomposed=True, n_units_affine=J["hidden"], non_lin_glow="relu",
                                 clamp_type="realnvp", flow_norm="actnorm", flow_batchnorm_momentum=0.0,
                                 learn_prior=True, n_units_prior=J["n_units_prior"], make_conditional=True,
                                 base_norm="actnorm", split2d_act="softplus", L=J["L"], K=J["K"], n_bits=J["n_bits"])


def cond_sizes(n):
    return [[n, c, 32 >> l, 32 >> l] for l, c in enumerate(J["cond_ch"])]


def trained_like(module, seed=0):
    """Random-init weights made non-trivial (zero-init Conv2dZeros / realnvp scale would make every
    coupling the identity); ActNorms marked initialised.  No checkpoint exists offline."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            p.add_(torch.randn(p.shape, generator=gen) * (0.01 if "conv.weight" in name else 0.05))
        for name, b in module.named_buffers():
            if name.endswith("initialized"):
                b.fill_(1)


def synth_inputs(n_frames, n_seq, seed):
    """Solver.preprocess-shaped data (RFN/trainer.py:165-175): floor(u*256)/256 - 0.5; SM-MNIST-like sparsity."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n_frames, 1, 64, 64, generator=g)
    mask = (torch.rand(n_frames, 1, 64, 64, generator=g) < 0.1).float()  # ~90 % black canvas
    x = torch.floor(u * mask * 256) / 256 - 0.5
    conds = [torch.randn(*s, generator=g) for s in cond_sizes(n_frames)]
    base = torch.randn(n_frames, J["base_ch"], 2, 2, generator=g)
    feats = torch.randn(n_seq, J["T"] - 1, J["lstm_in"], 2, 2, generator=g)
    return x, conds, base, feats


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference is pure Python and cannot travel to the GPU box)
# ------------------------------------------------------------------------------------------------
def cpu_step_builder(n_frames):
    import oracle as O
    import recurrent_flows_msc_b200 as rf
    torch.manual_seed(0)
    flow = rf.ListGlow([n_frames, 1, 64, 64], cond_sizes(n_frames), [n_frames, J["base_ch"], 2, 2], glow_args()).eval()
    trained_like(flow, 0)
    sd = {k: v.clone() for k, v in flow.state_dict().items()}
    lstm = rf.ConvLSTM(J["lstm_in"], J["lstm_hidden"], [3, 3])
    w, b = lstm.LSTMlayer.conv[0].weight.detach().clone(), lstm.LSTMlayer.conv[0].bias.detach().clone()
    x, conds, base, _ = synth_inputs(n_frames, 1, 1)
    feats = torch.randn(n_frames, 1, J["lstm_in"], 2, 2)
    noise = torch.rand(n_frames, 1, 64, 64) / 256

    def step():
        with torch.no_grad():
            O.convlstm(feats, w, b)
            _, nll = O.listglow_log_prob(x, conds, base, sd, J["L"], J["K"], J["n_bits"], noise=noise, learn_prior=True)
        return nll
    return step


def time_cpu_training(n_frames, steps=2, warmup=1):
    """CPU arm of the training step: torch autograd through the oracle port (forward + backward of ListGlow.log_prob and
    one ConvLSTM step, no optimizer) on a bounded sample of frames."""
    import oracle as O
    import recurrent_flows_msc_b200 as rf
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    flow = rf.ListGlow([n_frames, 1, 64, 64], cond_sizes(n_frames), [n_frames, J["base_ch"], 2, 2], glow_args())
    trained_like(flow, 0)
    leaf = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in flow.state_dict().items()}
    lstm = rf.ConvLSTM(J["lstm_in"], J["lstm_hidden"], [3, 3])
    w = lstm.LSTMlayer.conv[0].weight.detach().clone().requires_grad_()
    b = lstm.LSTMlayer.conv[0].bias.detach().clone().requires_grad_()
    x, conds, base, _ = synth_inputs(n_frames, 1, 1)
    feats = torch.randn(n_frames, 1, J["lstm_in"], 2, 2)
    noise = torch.rand(n_frames, 1, 64, 64) / 256

    def step():
        hs, _, _ = O.convlstm(feats, w, b)
        bc = torch.cat([hs[:, 0], base[:, J["lstm_hidden"]:]], 1)
        _, nll = O.listglow_log_prob(x, conds, bc, leaf, J["L"], J["K"], J["n_bits"], noise=noise, learn_prior=True)
        (nll.mean() / (math.log(2.0) * 4096)).backward()
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return n_frames / statistics.median(ts), statistics.median(ts), torch.get_num_threads()


def time_cpu(n_frames, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_builder(n_frames)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return n_frames / statistics.median(ts), statistics.median(ts), torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 64
    fps, t, cores = time_cpu(n, max(1, args.steps), max(0, args.warmup))
    sample = f"{n} frames per step: oracle ListGlow.log_prob (config J) + 1 ConvLSTM cell step at batch {n}, fp32, torch CPU"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "rfn_J_hotpath_fwd", "pass": "forward (density evaluation)", "frames_per_step": n},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [f.strip() for f in line.split(",")]))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.06] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class KernelTimer:
    """CUDA-event timing of every librfk launch on the launching (current) stream."""

    def __init__(self):
        self.rec = []
        self.elementwise = {}

    @contextlib.contextmanager
    def __call__(self, name, meta):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self.rec.append((name, meta, a, b))

    def table(self):
        torch.cuda.synchronize()
        agg, shapes = {}, {}
        for name, meta, a, b in self.rec:
            ms = a.elapsed_time(b)
            d = agg.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "flops_padded": 0.0})
            d["launches"] += 1
            d["ms"] += ms
            if meta and "flops" not in meta:   # bandwidth-bound kernel: algorithmic bytes only; keep its largest launch
                e = self.elementwise.setdefault(name, {})
                k = meta["bytes"]
                ee = e.setdefault(k, {"launches": 0, "ms": 0.0})
                ee["launches"] += 1
                ee["ms"] += ms
                continue
            if meta:
                d["flops"] += meta["flops"]
                d["flops_padded"] += meta["flops_padded"]
                k = (name, meta["M"], meta["N"], meta["K"])
                sdict = shapes.setdefault(k, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
                sdict["launches"] += 1
                sdict["ms"] += ms
                sdict["flops"] += meta["flops"]
                sdict["bytes"] += meta["bytes"]
        self.shapes = shapes
        return agg


def elementwise_at_scale(dev, hbm_peak):
    """Achieved algorithmic GB/s of the bandwidth-bound kernels on tensors larger than the 126 MB L2 (the flow tensors of
    config J are only 9 MB per level, so inside the workload these kernels are launch-latency-bound, not HBM-bound)."""
    import recurrent_flows_msc_b200 as rf
    ops = rf.ops
    out = []

    def timeit(name, nbytes, fn, iters=5):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        us = 1e3 * a.elapsed_time(b) / iters
        out.append({"kernel": name, "algorithmic_mb": round(nbytes / 1e6, 1), "us": round(us, 1),
                    "gbs": round(nbytes / us / 1e3, 1), "hbm_frac": round(nbytes / us / 1e3 / hbm_peak, 3)})

    B, C, H, W = 64, 48, 128, 128
    x = torch.randn(B, C, H, W, device=dev)
    n = x.numel()
    bias, logs = torch.randn(1, C, 1, 1, device=dev) * 0.1, torch.randn(1, C, 1, 1, device=dev) * 0.1
    timeit("rfk_squeeze2d", 8.0 * n, lambda: ops.squeeze2d(x, False))
    timeit("rfk_actnorm", 8.0 * n, lambda: ops.actnorm(x, bias, logs, False))
    for Cm in (4, 12, 48):   # the 1x1 mix runs on the fp32 pipes: HBM-bound only while 2*C flop/element stays under the ridge
        xm = x.view(B * C // Cm, Cm, H, W)
        Wm, bv = torch.randn(Cm, Cm, device=dev) / Cm ** 0.5, torch.randn(Cm, device=dev)
        timeit(f"rfk_mix1x1 (ActNorm+InvConv, C={Cm})", 8.0 * n, lambda: ops.mix1x1(xm, Wm, bv))
    nh = torch.empty(B, H, W, 64, device=dev, dtype=torch.bfloat16)
    timeit("rfk_pack_nhwc_bf16", 6.0 * n, lambda: ops.pack_nhwc(x, 0, C, nh, 0))
    ld = torch.zeros(B, device=dev)
    timeit("rfk_gauss_logp", 4.0 * n, lambda: ops.gauss_logp(x, 0, None, C, ops.PAIR_SPLIT, "exp", ld))
    del nh
    Bt, Ct = 16, 24
    taps = torch.randn(Bt, 9 * Ct, H, W, device=dev) * 0.1
    z = torch.randn(Bt, Ct, H, W, device=dev)
    sc, sh = torch.rand(Ct, device=dev) + 0.5, torch.randn(Ct, device=dev) * 0.1
    cs, csh = torch.randn(Ct // 2, device=dev) * 0.3, torch.randn(Ct // 2, device=dev) * 0.1
    ldt = torch.zeros(Bt, device=dev)
    timeit("rfk_coupling_tail_taps", 4.0 * taps.numel() + 4.0 * z.numel(),
           lambda: ops.coupling_tail_taps(taps, z, sc, sh, "realnvp", cs, csh, ldt, False))
    dz, gl = torch.randn_like(z), torch.randn(Bt, device=dev)
    timeit("rfk_coupling_taps_bwd", 4.0 * taps.numel() + 4.0 * z.numel() * 2.5,
           lambda: ops.coupling_taps_bwd(taps, z, dz, sc, sh, "realnvp", cs, csh, gl, 3.0))
    del taps, z, dz
    Hc = 64
    cc = torch.randn(32, 4 * Hc, 64, 64, device=dev)
    cp = torch.randn(32, Hc, 64, 64, device=dev)
    timeit("rfk_convlstm_pointwise", 4.0 * cp.numel() * 7, lambda: ops.convlstm_pointwise(cc, cp, None))
    dhh = torch.randn(32, Hc, 64, 64, device=dev)
    timeit("rfk_convlstm_pointwise_bwd", 4.0 * cp.numel() * 11, lambda: ops.convlstm_pointwise_bwd(cc, cp, None, dhh, None, None))
    del cc, cp, dhh
    rows = 583680
    dh = torch.randn(rows, 256, device=dev).to(torch.bfloat16)
    hh = torch.randn(rows, 256, device=dev).to(torch.bfloat16)
    s256 = torch.rand(256, device=dev) + 0.5
    timeit("rfk_act_affine_bwd", 6.0 * rows * 256, lambda: ops.act_affine_bwd(dh, hh, 256, s256, "relu"))
    del dh, hh
    npar = 1 << 26
    p4 = [torch.zeros(npar, device=dev) for _ in range(4)]
    st = torch.ones(1, device=dev)
    timeit("rfk_adam_step", 28.0 * npar,
           lambda: rf._lib.call("rfk_adam_step", p4[0].data_ptr(), p4[1].data_ptr(), p4[2].data_ptr(), p4[3].data_ptr(), npar,
                                1e-3, 0.9, 0.999, 1e-8, 1.0, st.data_ptr(), ops._stream()))
    return out


def run_ours(args):
    import torch.distributed as dist
    import recurrent_flows_msc_b200 as rf

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's version banner goes to stdout and would break the one-JSON-line contract
        os.environ["NCCL_DEBUG"] = os.environ.get("RFK_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    B, T = J["B"], J["T"]
    n_frames = B * (T - 1)

    # identical replicas (same seed); each rank draws its own shard of sequences
    torch.manual_seed(0)
    flow = rf.ListGlow([n_frames, 1, 64, 64], cond_sizes(n_frames), [n_frames, J["base_ch"], 2, 2], glow_args()).eval()
    trained_like(flow, 0)
    lstm = rf.ConvLSTM(J["lstm_in"], J["lstm_hidden"], [3, 3]).eval()
    flow, lstm = flow.to(dev), lstm.to(dev)
    hx, hconds, hbase, hfeats = synth_inputs(n_frames, B, 1 + rank)
    host = [t.pin_memory() for t in [hx, hbase, hfeats] + hconds]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)
    nll_host = torch.empty(n_frames, dtype=torch.float32).pin_memory()
    h_host = torch.empty(B, J["lstm_hidden"], 2, 2, dtype=torch.float32).pin_memory()
    d2h_bytes = nll_host.numel() * 4 + h_host.numel() * 4

    def upload():
        dx, dbase, dfeats, *dconds = [t.to(dev, non_blocking=True) for t in host]
        return dx, dconds, dbase, dfeats

    def hot_path(dx, dconds, dbase, dfeats):
        with torch.no_grad():
            _, h_last, _ = lstm(dfeats)
            _, nll = flow.log_prob(dx, dconds, dbase)
        return nll, h_last

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    resident = upload()
    eager_hot_path = hot_path
    for _ in range(max(args.warmup, 3)):
        nll, _ = hot_path(*resident)
    torch.cuda.synchronize()
    assert torch.isfinite(nll).all(), "non-finite nll in warm-up"
    launches_per_step = rf._lib.launches
    eager_hot_path(*resident)
    launches_per_step = rf._lib.launches - launches_per_step
    if not args.no_graph:
        # the same launches, replayed from a CUDA graph: removes ~20 us of Python/ctypes per launch from the host side
        graphed = rf.Graphed(lambda x, c, b, f: eager_hot_path(x, c, b, f), resident[0], resident[1], resident[2], resident[3])
        hot_path = lambda x, c, b, f: graphed(x, c, b, f)  # noqa: E731
        for _ in range(3):
            nll, _ = hot_path(*resident)
        torch.cuda.synchronize()
        assert torch.isfinite(nll).all(), "non-finite nll in graph replay"

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- device-resident timing ---------------------------------------------------------------
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        hot_path(*resident)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = launches_per_step * args.steps   # graph replays do not pass through the ctypes counter
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    # ---- end-to-end from pinned host buffers ------------------------------------------------------
    # Every step: H2D of that step's inputs (pinned host -> device), the hot path, D2H of nll and h.  With graphs the
    # copies run on a second stream into the other of two graph instances' static inputs, so the transfer of step
    # i+1 overlaps the compute of step i (what a double-buffered data loader does); eager mode keeps it serial.
    hx_, hbase_, hfeats_, *hconds_ = host
    if args.no_graph:
        def e2e_loop(n):
            for _ in range(n):
                nll, h_last = hot_path(*upload())
                nll_host.copy_(nll, non_blocking=True)
                h_host.copy_(h_last, non_blocking=True)
    else:
        graphs = [graphed, rf.Graphed(lambda x, c, b, f: eager_hot_path(x, c, b, f), resident[0], resident[1], resident[2],
                                      resident[3])]
        copy_stream = torch.cuda.Stream()

        def e2e_loop(n):
            main = torch.cuda.current_stream()
            copied = [torch.cuda.Event() for _ in range(n)]
            done = [torch.cuda.Event() for _ in range(n)]
            for i in range(n):
                gi = graphs[i % 2]
                with torch.cuda.stream(copy_stream):
                    if i >= 2:
                        copy_stream.wait_event(done[i - 2])        # this instance's inputs are free again
                    else:
                        copy_stream.wait_stream(main)
                    sx, sc, sb, sf = gi.static_in
                    sx.copy_(hx_, non_blocking=True)
                    sb.copy_(hbase_, non_blocking=True)
                    sf.copy_(hfeats_, non_blocking=True)
                    for d_, h_ in zip(sc, hconds_):
                        d_.copy_(h_, non_blocking=True)
                    copied[i].record(copy_stream)
                main.wait_event(copied[i])
                gi.graph.replay()
                nll, h_last = gi.out
                nll_host.copy_(nll, non_blocking=True)
                h_host.copy_(h_last, non_blocking=True)
                done[i].record(main)

    e2e_loop(2)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    e2e_loop(args.steps)
    ev3.record()
    barrier()
    ms_e2e = torch.tensor([ev2.elapsed_time(ev3)], device=dev)
    # ---- sampling direction (RFN.predict's inner step): ConvLSTM cell + ListGlow.sample for B sequences -------
    sconds = [c[:B].contiguous() for c in resident[1]]
    sbase = resident[2][:B].contiguous()
    sfeat = resident[3][:, :1].contiguous()

    def sample_step_eager():
        with torch.no_grad():
            lstm(sfeat)
            return flow.sample(None, sconds, sbase, num_samples=B, temperature=0.7)

    def time_loop(fn, n):
        for _ in range(3):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / n], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    n_samp = max(args.steps, 10)
    ms_sample = time_loop(sample_step_eager, n_samp)
    gsample = rf.GraphedSample(flow, sconds, sbase, temperature=0.7)
    ms_sample_graph = time_loop(lambda: gsample(sconds, sbase), n_samp)
    assert torch.isfinite(gsample(sconds, sbase)).all()
    # ---- training step of the flow decoder: forward with tape + hand-written backward + gradient allreduce + fused Adam
    training = None
    if not args.no_train:
        import copy
        tflow, tlstm = copy.deepcopy(flow).train(), copy.deepcopy(lstm).train()
        opt = rf.FlatAdam(list(tflow.parameters()) + list(tlstm.parameters()), lr=1e-4, world_size=world)
        tx, tconds, tfeats = resident[0], resident[1], resident[3]
        hcn = J["lstm_hidden"]
        z_part = resident[2][:, hcn:].contiguous()   # the latent-sample part of RFN's base condition cat[h_t, z_t]

        def loss_fn():
            # as in RFN.loss: the recurrence's hidden states condition the flow's prior, so the flow's gradient w.r.t. its
            # base condition is back-propagated through all 19 ConvLSTM steps
            hs, _, _ = tlstm(tfeats)
            tbase = torch.cat([hs.reshape(n_frames, hcn, 2, 2), z_part], 1)
            _, nll = tflow.log_prob(tx, tconds, tbase)
            return nll.mean() / (math.log(2.0) * 64 * 64)

        l0 = rf._lib.launches
        opt.zero_grad()
        first_loss = loss_fn()
        first_loss.backward()
        opt.step()
        first_loss = float(first_loss.detach())
        train_launches = rf._lib.launches - l0
        del l0
        n_train = max(5, args.steps // 2)
        if args.no_graph:
            def train_step():
                opt.zero_grad()
                loss = loss_fn()
                loss.backward()
                opt.step()
                return loss.detach()
        else:
            train_step = rf.GraphedTrainStep(loss_fn, opt, warmup=2)
        ms_train = time_loop(train_step, n_train)
        last_loss = float(train_step())
        assert math.isfinite(last_loss), "non-finite training loss"
        training = {"what": "hot-path training step on 570 frames per GPU: 19 ConvLSTM steps whose hidden states form the flow's "
                            "base condition, ListGlow.log_prob forward recording a tape, backward on hand-written kernels "
                            "(coupling / ActNorm / Split2d / prior backward, tcgen05 wgrad + dgrad, BPTT through the ConvLSTM), "
                            "one NCCL sum-allreduce of the flat gradient, Adam for all parameters in one launch, bf16 weight "
                            "repacking; inputs resident in HBM",
                    "frames_per_s": world * n_frames / (ms_train / 1e3), "ms_per_step": ms_train, "steps": n_train,
                    "launch": "eager" if args.no_graph else "two CUDA graphs (fwd+bwd+gather | Adam) around the eager allreduce",
                    "own_kernel_launches_per_step": train_launches, "parameters": opt.n,
                    "allreduce_bytes_per_step": opt.n_pad * 4 if world > 1 else 0,
                    "bits_per_dim_first_step": first_loss, "bits_per_dim_last_step": last_loss,
                    "peak_memory_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        del train_step, opt, tflow, tlstm
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.summary(t_wall0, t_wall1) if sampler else None

    if rank == 0:
        ms_step = float(ms) / args.steps
        ms_step_e2e = float(ms_e2e) / args.steps
        value = world * n_frames / (ms_step / 1e3)
        e2e = world * n_frames / (ms_step_e2e / 1e3)
        # ---- per-kernel CUDA-event pass (same step, outside the timed regions) -> roofline of the dominant kernel
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        eager_hot_path(*resident)     # the training section invalidated the weight caches: rebuild them outside the traced step
        torch.cuda.synchronize()
        kt = KernelTimer()
        rf._lib.tracer = kt
        torch.cuda.profiler.start()   # `ncu --profile-from-start off` captures exactly this one step
        eager_hot_path(*resident)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        rf._lib.tracer = None
        table = kt.table()
        total_ms = sum(d["ms"] for d in table.values())
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        by_shape = []
        for (name, M, N, K), v in sorted(kt.shapes.items(), key=lambda kv: -kv[1]["ms"])[:6]:
            sec = v["ms"] / 1e3
            by_shape.append({"kernel": name, "M": M, "N": N, "K": K, "launches": v["launches"],
                             "avg_us": round(1e3 * v["ms"] / v["launches"], 1),
                             "tflops": round(v["flops"] / sec / 1e12, 1), "tensor_frac": round(v["flops"] / sec / 1e12 / peak_tf, 3),
                             "algorithmic_gbs": round(v["bytes"] / sec / 1e9, 1), "hbm_frac": round(v["bytes"] / sec / 1e9 / hbm_peak, 3)})
        # dominant kernel = the (entry point, GEMM shape) with the largest share of the step; its binding roofline is the
        # one it sits closer to (these N <= 256 layers straddle the ridge: AI ~ 100-250 flop/B)
        top = by_shape[0]
        (tk, tv) = max(kt.shapes.items(), key=lambda kv: kv[1]["ms"])
        sec = tv["ms"] / 1e3
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            hit = tj.get("|".join(str(v) for v in tk))
            traffic = hit["traffic_bytes"] if hit else None
        except (OSError, ValueError):
            pass
        bound = "hbm" if top["hbm_frac"] >= top["tensor_frac"] else "tensor"
        roofline = {"kernel": f"{tk[0]} M={tk[1]} N={tk[2]} K={tk[3]}", "bound": bound,
                    "achieved": top["algorithmic_gbs"] if bound == "hbm" else top["tflops"],
                    "peak": hbm_peak if bound == "hbm" else peak_tf, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                    "frac": top["hbm_frac"] if bound == "hbm" else top["tensor_frac"], "traffic": traffic,
                    "peak_source": ("MEASURED_PEAKS.json (hbm_gbs; bf16_tflops_sustained: kernel timed inside a long step)"
                                    if peaks else "fallback 6.65 TB/s, 1.4 PFLOP/s sustained (B200_PROFILING.md)"),
                    "algorithmic_bytes_per_launch": tv["bytes"] / tv["launches"],
                    "algorithmic_flops_per_launch": tv["flops"] / tv["launches"],
                    "share_of_step": tv["ms"] / total_ms, "launches_per_step": tv["launches"],
                    "avg_launch_us": top["avg_us"], "tensor_frac": top["tensor_frac"], "hbm_frac": top["hbm_frac"],
                    "by_shape": by_shape,
                    "note": "per-launch figures from CUDA events on the launching stream in an instrumented extra step; "
                            "algorithmic bytes = activations in + out + weights"}
        # bandwidth-bound kernels: achieved algorithmic GB/s of each kernel's LARGEST launch shape (level 1) vs the HBM peak
        elementwise = []
        for name, by_bytes in kt.elementwise.items():
            nbytes = max(by_bytes)
            v = by_bytes[nbytes]
            us = 1e3 * v["ms"] / v["launches"]
            elementwise.append({"kernel": name, "algorithmic_bytes_per_launch": nbytes, "launches": v["launches"],
                                "avg_us": round(us, 1), "gbs": round(nbytes / us / 1e3, 1),
                                "hbm_frac": round(nbytes / us / 1e3 / hbm_peak, 3)})
        elementwise.sort(key=lambda e: -e["algorithmic_bytes_per_launch"])
        roofline["elementwise_in_workload_largest_shape"] = elementwise
        roofline["elementwise_at_scale"] = elementwise_at_scale(dev, hbm_peak)
        kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 4),
                       "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["flops"] else None}
                   for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"])}
        n_cpu = 64
        cpu_fps, cpu_t, cores = time_cpu(n_cpu, 3, 1)
        if training is not None:
            n_cpu_t = 32
            tfps, tt, tcores = time_cpu_training(n_cpu_t)
            training["cpu_baseline"] = {"value": tfps, "unit": "frames/s", "cores": tcores, "kind": "port",
                                        "sample": f"{n_cpu_t} frames: torch CPU autograd through the oracle port (ListGlow.log_prob "
                                                  f"config J + 1 ConvLSTM step, forward + backward, no optimizer), median of 2 "
                                                  f"after 1 warm-up ({tt:.2f} s each)"}
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "rfn_J_hotpath_fwd",
                       "pass": "forward (density evaluation): 19 ConvLSTM steps + ListGlow.log_prob on 570 frames; the training step "
                               "(forward + hand-written backward + all-reduce + Adam) is reported under `training`",
                       "frames_per_step_per_gpu": n_frames, "sequences_per_gpu": B, "L": J["L"], "K": J["K"],
                       "hidden": J["hidden"], "conv_dtype": "bf16 in / fp32 accumulate", "flow_dtype": "f32",
                       "l2": "inputs_exceed_L2 (>=300 MB of activations per level-1 GlowStep vs 126 MB L2)",
                       "parallelism": f"batch-sharded x{world}, no data-path collective",
                       "launch": "eager (Python/ctypes per launch)" if args.no_graph else "CUDA graph replay of the same launches"},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": ms_step_e2e,
                    "how": "serial H2D -> compute -> D2H" if args.no_graph else
                           "H2D of step i+1 on a copy stream overlaps the graph replay of step i (two graph instances)"},
            "sampling": {"what": "one RFN.predict inner step: ListGlow.sample (reverse flow, T=0.7) for 30 sequences "
                                 "(+ ConvLSTM cell in the eager figure); autoregressive, so only the batch is parallel",
                         "eager_frames_per_s": world * B / (ms_sample / 1e3), "eager_ms": ms_sample,
                         "cuda_graph_frames_per_s": world * B / (ms_sample_graph / 1e3), "cuda_graph_ms": ms_sample_graph},
            "training": training,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels_ms_per_step": kernels,
            "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n_cpu} frames: oracle ListGlow.log_prob (config J) + 1 ConvLSTM step at batch {n_cpu}, "
                                       f"median of 3 after 1 warm-up ({cpu_t:.2f} s each)"},
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-train", action="store_true", help="skip the training-step section")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every launch from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
