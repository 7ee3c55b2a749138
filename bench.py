#!/usr/bin/env python
"""Benchmark of the RFN hot path (Glow decoder + ConvLSTM recurrence) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

The metric of record (BASELINE.json: "RFN 64x64 train/sample frames/sec") is the TRAINING step of the reference's
job-script configuration J (RFN/default_rfn_job.sh: B=30 sequences per GPU, 1x64x64, 10+10 frames, L=5, K=10, hidden
256, h=200, z=56), workload ``rfn_J_train``:

  forward   19 ConvLSTM cell steps (512 -> 200 hidden channels, 3x3, 2x2 maps, batch 30) whose hidden states form the
            flow's base condition (as in RFN.loss, RFN/RFN_new.py:131-139,196), then ListGlow.log_prob (dequantise, f: 5
            levels x 10 GlowSteps + 4 Split2d, learned prior) on the B*(T-1) = 570 predicted frames, time-batched into one
            call (exact because nothing the flow produces feeds back into the recurrence)
  backward  hand-written kernels over the recorded tape, incl. BPTT through the ConvLSTM
  update    NCCL sum all-reduce of the flat gradient (inside the timed step for N > 1) + Adam for all parameters
            (RFN/trainer.py:241-248)

``value`` = frames/s of that step with the batch resident in HBM (CUDA-graph replay; max over ranks); ``e2e`` = the same
step fed from pinned HOST buffers (H2D of x, the condition pyramid, z_t and the ConvLSTM inputs; D2H of the loss) with
the copies inside the timed region.  The same line carries the forward-only (density evaluation) and sampling
(RFN.predict inner step) throughput of the same models under ``forward`` / ``sampling``, the roofline of the dominant
kernel family of the training step, a parity block (full-depth J sub-batch vs the CPU oracle) and the CPU baseline.

Other workloads (--workload): rfn_J_fwd, rfn_J_sample, glow_cfg1, convlstm_cfg2, rfn_J_smooth_D3, rfn_D_sample
(BASELINE.md section 3 configs 1, 2, 4, 5).

With --gpus N (torchrun) every rank processes its own batch of sequences (weak scaling); time = max over ranks.
``--impl reference`` times the reference's own modules (baseline/_ref, staged by __graft_entry__.build(); the oracle port
when that is absent) on the host cores, same pass, same frames, rank 0 only.
"""
import argparse
import contextlib
import copy
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

if "--impl" in sys.argv and "reference" in sys.argv:
    # the CPU arm: the reference binds `device = cuda if available` at import (Utils/modules.py:4, Flow/glow.py:8)
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

import torch  # noqa: E402

METRIC = "RFN 64x64 train frames/sec"
J = dict(name="J", B=30, T=20, C=1, L=5, K=10, hidden=256, n_units_prior=512, cond_ch=[16, 32, 64, 128, 256],
         base_ch=256, lstm_in=512, lstm_hidden=200, z_dim=56, n_bits=8)
# main_rfn.py defaults (config D): 3x64x64, L5 K15, with_skip condition channels, h=256, z=5, extractor features 256
D = dict(name="D", B=32, T=30, C=3, L=5, K=15, hidden=256, n_units_prior=512, cond_ch=[32, 64, 128, 256, 384],
         base_ch=261, lstm_in=256, lstm_hidden=256, z_dim=5, n_bits=8)
WORKLOADS = ("rfn_J_train", "rfn_J_fwd", "rfn_J_sample", "glow_cfg1", "convlstm_cfg2", "rfn_J_smooth_D3", "rfn_D_sample")
METRICS = {"rfn_J_train": METRIC, "rfn_J_smooth_D3": "RFN 64x64 train frames/sec (smoothing + overshooting D=3 hot path)",
           "rfn_J_fwd": "RFN 64x64 density-evaluation frames/sec", "rfn_J_sample": "RFN 64x64 sample frames/sec",
           "rfn_D_sample": "RFN-VGG-Glow 3x64x64 sample frames/sec (2 context + 28 predicted)",
           "glow_cfg1": "Glow L3 K8 forward images/sec", "convlstm_cfg2": "ConvLSTM 64ch 64x64 T=10 frames/sec"}


def glow_args(cfg):
    return types.SimpleNamespace(LU_decomposed=True, n_units_affine=cfg["hidden"], non_lin_glow="relu",
                                 clamp_type="realnvp", flow_norm="actnorm", flow_batchnorm_momentum=0.0,
                                 learn_prior=True, n_units_prior=cfg["n_units_prior"], make_conditional=True,
                                 base_norm="actnorm", split2d_act="softplus", L=cfg["L"], K=cfg["K"], n_bits=cfg["n_bits"])


def cond_sizes(cfg, n):
    return [[n, c, 32 >> l, 32 >> l] for l, c in enumerate(cfg["cond_ch"])]


def trained_like(module, seed=0):
    """Random-init weights made non-trivial (zero-init Conv2dZeros / realnvp scale would make every
    coupling the identity); ActNorms marked initialised.  No checkpoint exists offline."""
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            # small enough that the 50..75-step flow stays in a trained model's regime (bits/dim O(10), samples within the image
            # range); 2.5x larger perturbations compound to > 1e3 bits/dim at K=10 and ~1e26 at K=15
            p.add_(torch.randn(p.shape, generator=gen) * (0.004 if "conv.weight" in name else 0.02))
        for name, b in module.named_buffers():
            if name.endswith("initialized"):
                b.fill_(1)


def synth_inputs(cfg, n_frames, n_seq, seed, dense=False):
    """Solver.preprocess-shaped data (RFN/trainer.py:165-175): floor(u*256)/256 - 0.5; SM-MNIST-like sparsity
    (~90 % black canvas) or, with dense=True, KTH/BAIR-like dense frames."""
    g = torch.Generator().manual_seed(seed)
    C = cfg["C"]
    u = torch.rand(n_frames, C, 64, 64, generator=g)
    if not dense:
        u = u * (torch.rand(n_frames, 1, 64, 64, generator=g) < 0.1).float()
    x = torch.floor(u * 256) / 256 - 0.5
    conds = [torch.randn(*s, generator=g) for s in cond_sizes(cfg, n_frames)]
    z_part = torch.randn(n_frames, cfg["z_dim"], 2, 2, generator=g)     # the latent sample z_t of cat[h_t, z_t]
    feats = torch.randn(n_seq, n_frames // n_seq, cfg["lstm_in"], 2, 2, generator=g)
    return x, conds, z_part, feats


def build_models(cfg, n_frames, mod=None):
    """(flow, lstm) of configuration cfg with trained-like weights, on the CPU.  `mod` = package providing ListGlow /
    ConvLSTM (ours by default); identical seeds give identical weights whichever package builds them."""
    if mod is None:
        import recurrent_flows_msc_b200 as mod
    torch.manual_seed(0)
    flow = mod.ListGlow([n_frames, cfg["C"], 64, 64], cond_sizes(cfg, n_frames), [n_frames, cfg["base_ch"], 2, 2], glow_args(cfg))
    trained_like(flow, 0)
    lstm = mod.ConvLSTM(cfg["lstm_in"], cfg["lstm_hidden"], [3, 3])
    return flow, lstm


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own modules from baseline/_ref (kind "reference"), else the oracle port (kind "port")
# ------------------------------------------------------------------------------------------------
class _PortFlow:
    """Oracle-port stand-in with the two calls the CPU arm makes (only when baseline/_ref is absent)."""

    def __init__(self, sd, cfg):
        self.cfg = cfg
        self.sd = {k: (v.clone().requires_grad_() if v.is_floating_point() else v.clone()) for k, v in sd.items()}

    def parameters(self):
        return [v for v in self.sd.values() if v.requires_grad]

    def log_prob(self, x, conds, base, logdet=0):
        import oracle as O
        noise = torch.rand_like(x) / 2 ** self.cfg["n_bits"]
        return O.listglow_log_prob(x, conds, base, self.sd, self.cfg["L"], self.cfg["K"], self.cfg["n_bits"], noise=noise,
                                   learn_prior=True)

    def sample(self, z, conds, base, temperature=0.7):
        import oracle as O
        cfg = self.cfg
        n = base.shape[0]
        cz = cfg["C"] * 2 ** (cfg["L"] + 1)
        eps_prior = torch.randn(n, cz, 2, 2)
        eps = [torch.randn(n, (cfg["C"] * 2) << l, 32 >> l, 32 >> l) for l in range(cfg["L"] - 1)]
        with torch.no_grad():
            return O.listglow_sample(conds, base, self.sd, cfg["L"], cfg["K"], eps_prior, eps, temperature, learn_prior=True)


class _PortLSTM:
    def __init__(self, w, b):
        self.w, self.b = w.clone().requires_grad_(), b.clone().requires_grad_()

    def parameters(self):
        return [self.w, self.b]

    def __call__(self, x, ht=None, ct=None):
        import oracle as O
        return O.convlstm(x, self.w, self.b, ht, ct)


def reference_models(cfg, n_frames):
    """The reference's ListGlow / ConvLSTM with OUR arm's weights (same seeds -> same state_dict), on the CPU."""
    flow0, lstm0 = build_models(cfg, n_frames)
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "Flow")):
        sys.path.insert(0, ref_dir)
        import warnings
        warnings.filterwarnings("ignore")
        from Flow import ListGlow as RefGlow           # the unmodified reference (Flow/glow.py:43)
        from Utils import ConvLSTM as RefLSTM          # Utils/modules.py:396
        torch.manual_seed(0)
        flow = RefGlow([n_frames, cfg["C"], 64, 64], cond_sizes(cfg, n_frames), [n_frames, cfg["base_ch"], 2, 2], glow_args(cfg))
        flow.load_state_dict(flow0.state_dict())
        lstm = RefLSTM(cfg["lstm_in"], cfg["lstm_hidden"], [3, 3], bias=True, peephole=True)
        lstm.load_state_dict(lstm0.state_dict())
        return flow, lstm, "reference"
    conv = lstm0.LSTMlayer.conv[0]
    return _PortFlow(flow0.state_dict(), cfg), _PortLSTM(conv.weight.detach(), conv.bias.detach()), "port"


def reference_step_builder(workload, n_frames=None):
    """Returns (step_fn, units_per_step, kind, sample_text) for the CPU arm of `workload`."""
    torch.set_num_threads(os.cpu_count() or 1)
    if workload in ("rfn_J_train", "rfn_J_fwd", "rfn_J_smooth_D3"):
        cfg = J
        B, T = cfg["B"], cfg["T"]
        # bounded sample (default 5 of the 19 time steps = 150 frames: every time step is the same call at the same shapes,
        # so frames/s does not depend on how many are timed; --ref-frames 570 runs them all)
        steps_t = min(T - 1, 5 if n_frames is None else max(1, n_frames // B))
        nf = B * steps_t
        flow, lstm, kind = reference_models(cfg, B)
        x, conds, z_part, feats = synth_inputs(cfg, nf, B, 1, dense=workload == "rfn_J_smooth_D3")
        xs = x.view(B, steps_t, *x.shape[1:])
        smooth = None
        if workload == "rfn_J_smooth_D3":
            torch.manual_seed(1)
            smooth = type(lstm)(cfg["lstm_in"] + cfg["lstm_hidden"], cfg["lstm_hidden"], [3, 3]) if kind == "reference" else None
            if smooth is None:
                import recurrent_flows_msc_b200 as rf
                conv = rf.ConvLSTM(cfg["lstm_in"] + cfg["lstm_hidden"], cfg["lstm_hidden"], [3, 3]).LSTMlayer.conv[0]
                smooth = _PortLSTM(conv.weight.detach(), conv.bias.detach())
        train = workload != "rfn_J_fwd"
        params = list(flow.parameters()) + list(lstm.parameters()) + (list(smooth.parameters()) if smooth else [])
        opt = torch.optim.Adam(params, lr=1e-4) if train else None
        if hasattr(flow, "train"):
            flow.train(train)

        def step():
            # the reference's call pattern (RFN/RFN_new.py:131-139,158-211): one ConvLSTM cell and one flow.log_prob per
            # time step at batch B.  The flow activations of a time step are released by an immediate backward of that
            # step's loss term (gradient accumulation, same sum) so that 570 frames fit in host memory.
            with torch.set_grad_enabled(train):
                if train:
                    opt.zero_grad()
                hs, h, c = [], None, None
                for t in range(steps_t):
                    _, h, c = lstm(feats[:, t:t + 1], h, c)
                    hs.append(h)
                if smooth is not None:
                    a = ca = None
                    for t in reversed(range(steps_t)):
                        _, a, ca = smooth(torch.cat([hs[t], feats[:, t]], 1).unsqueeze(1), a, ca)
                        hs[t] = hs[t] + 0.0 * a.mean()      # the smoothing state reaches the loss through the encoder (f2, not timed)
                total = 0.0
                for t in range(steps_t):
                    base = torch.cat([hs[t], z_part.view(B, steps_t, -1, 2, 2)[:, t]], 1)
                    ct = [cc.view(B, steps_t, *cc.shape[1:])[:, t] for cc in conds]
                    _, nll = flow.log_prob(xs[:, t], ct, base, 0)
                    loss = nll.sum() / (nf * math.log(2.0) * 64 * 64)
                    if train:
                        loss.backward(retain_graph=t + 1 < steps_t)
                    total += float(loss.detach())
                if train:
                    opt.step()
            return total
        what = ("training step (forward + autograd backward + Adam)" if train else "forward (density evaluation)")
        sample = (f"{nf} frames per step ({steps_t} of the {T - 1} time steps at batch {B}): {what} of the reference's ListGlow.log_prob "
                  f"(config J) + ConvLSTM, one call per time step as RFN.loss does, fp32, torch CPU")
        return step, nf, kind, sample
    if workload in ("rfn_J_sample", "rfn_D_sample"):
        cfg = J if workload == "rfn_J_sample" else D
        B = cfg["B"] if n_frames is None else min(cfg["B"], n_frames)
        n_pred = 1 if workload == "rfn_J_sample" else (28 if n_frames is None else 2)
        flow, lstm, kind = reference_models(cfg, B)
        _, conds, z_part, feats = synth_inputs(cfg, B, B, 1)
        if hasattr(flow, "eval"):
            flow.eval()

        def step():
            h = c = None
            with torch.no_grad():
                for _ in range(n_pred):
                    _, h, c = lstm(feats[:, :1], h, c)
                    xs = flow.sample(None, conds, torch.cat([h, z_part], 1), temperature=0.7)
            return float(xs.mean())
        return step, B * n_pred, kind, (f"{B} sequences x {n_pred} predicted frames: reference ConvLSTM cell + ListGlow.sample "
                                        f"(config {cfg['name']}) per frame, fp32, torch CPU")
    if workload == "glow_cfg1":
        flow, sd, x, conds = cfg1_model()
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        kind = "port"
        if os.path.isdir(os.path.join(ref_dir, "Flow")):
            sys.path.insert(0, ref_dir)
            from Flow import ListGlow as RefGlow
            ref = RefGlow([16, 1, 32, 32], [list(c.shape) for c in conds], [16, 0, 4, 4], cfg1_args()).eval()
            ref.load_state_dict(sd)
            kind = "reference"

            def step():
                with torch.no_grad():
                    return float(ref.log_prob(x, conds, None, 0)[1].mean())
        else:
            def step():
                import oracle as O
                with torch.no_grad():
                    return float(O.listglow_log_prob(x, conds, None, sd, 3, 8, 8, noise=None, learn_prior=False,
                                                     make_conditional=False)[1].mean())
        return step, 16, kind, "16 images: ListGlow.log_prob L=3 K=8 (cfg1), fp32, torch CPU"
    if workload == "convlstm_cfg2":
        Bc = 32 if n_frames is None else max(1, n_frames // 10)
        import recurrent_flows_msc_b200 as rf
        torch.manual_seed(0)
        l0 = rf.ConvLSTM(64, 64, [3, 3])
        _, lstm, kind = None, None, "port"
        ref_dir = os.path.join(ROOT, "baseline", "_ref")
        if os.path.isdir(os.path.join(ref_dir, "Utils")):
            sys.path.insert(0, ref_dir)
            from Utils import ConvLSTM as RefLSTM
            lstm = RefLSTM(64, 64, [3, 3], bias=True, peephole=True)
            lstm.load_state_dict(l0.state_dict())
            kind = "reference"
        else:
            conv = l0.LSTMlayer.conv[0]
            lstm = _PortLSTM(conv.weight.detach(), conv.bias.detach())
        xs = torch.randn(Bc, 10, 64, 64, 64)

        def step():
            with torch.no_grad():
                return float(lstm(xs)[1].mean())
        return step, Bc * 10, kind, f"{Bc} sequences x 10 steps: ConvLSTM 64->64, 3x3, 64x64 maps (cfg2), fp32, torch CPU"
    raise SystemExit(f"unknown workload {workload}")


def cfg1_args():
    return types.SimpleNamespace(LU_decomposed=True, n_units_affine=256, non_lin_glow="relu", clamp_type="realnvp",
                                 flow_norm="actnorm", flow_batchnorm_momentum=0.0, learn_prior=False, n_units_prior=512,
                                 make_conditional=False, base_norm="actnorm", split2d_act="softplus", L=3, K=8, n_bits=8)


def cfg1_model():
    """BASELINE config 1: Glow L=3 K=8 hidden 256 on a 1x32x32 batch of 16, unconditional, N(0,1) prior."""
    import recurrent_flows_msc_b200 as rf
    B = 16
    sizes = [[B, 0, 16, 16], [B, 0, 8, 8], [B, 0, 4, 4]]
    torch.manual_seed(0)
    flow = rf.ListGlow([B, 1, 32, 32], sizes, [B, 0, 4, 4], cfg1_args()).eval()
    trained_like(flow, 1)
    g = torch.Generator().manual_seed(0)
    x = torch.floor(torch.rand(B, 1, 32, 32, generator=g) * 256) / 256 - 0.5
    return flow, {k: v.clone() for k, v in flow.state_dict().items()}, x, [torch.zeros(*s) for s in sizes]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, units, kind, sample = reference_step_builder(args.workload, args.ref_frames)
    for _ in range(max(0, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps)):
        step()
    t = (time.perf_counter() - t0) / max(1, args.steps)
    fps = units / t
    cores = torch.get_num_threads()
    print(json.dumps({
        "impl": "reference", "metric": METRICS[args.workload], "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, int(os.environ.get("WORLD_SIZE", "1")), args),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_subprocess(workload, ref_frames):
    """cpu_baseline leg of the GPU arm: the reference arm on a bounded sample, in a CUDA-less child process."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "1",
           "--warmup", "1", "--ref-frames", str(ref_frames)]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": f"failed: {e!r}"[:300]}


def workload_config(workload, world, args):
    """Identical for both arms: the reference arm prints OUR arm's config (it times a bounded sample of that workload on
    the host cores, stated in its cpu_baseline.sample; the dtype / launch keys describe the GPU arm)."""
    base = _workload_config(workload)
    base.update(conv_dtype="bf16 operands / fp32 accumulate (tcgen05 kind::f16)", flow_dtype="f32",
                l2=("inputs_exceed_L2 (level-1 hidden activations are 2 x 299 MB per GlowStep vs 126 MB L2)"
                    if workload.startswith("rfn_J") and "sample" not in workload else "working set below L2: latency-bound"),
                parallelism=(f"batch-sharded x{world}; one NCCL sum all-reduce of the flat fp32 gradient per step"
                             if "train" in workload or "smooth" in workload else f"batch-sharded x{world}, no data-path collective"),
                launch="eager (Python/ctypes per launch)" if args.no_graph else "CUDA graph replay")
    return base


def _workload_config(workload):
    base = {"workload": workload}
    if workload.startswith("rfn_J"):
        base.update({"config": "J (RFN/default_rfn_job.sh)", "sequences_per_gpu": J["B"], "frames_per_step_per_gpu": J["B"] * (J["T"] - 1),
                     "frames": "1x64x64, 10 conditioning + 10 predicted", "L": J["L"], "K": J["K"], "hidden": J["hidden"]})
    if workload in ("rfn_J_train", "rfn_J_smooth_D3"):
        base["pass"] = ("training step: forward (19 ConvLSTM steps feeding the flow's base condition + time-batched ListGlow.log_prob) + "
                        "backward + gradient all-reduce + Adam")
    if workload == "rfn_J_smooth_D3":
        base["extra"] = "second (smoothing) ConvLSTM 712 -> 200 over 19 steps, dense KTH-shaped frames; overshooting D=3 re-rolls only the prior net (outside the hot path)"
    if workload == "rfn_J_fwd":
        base["pass"] = "forward (density evaluation)"
    if workload == "rfn_J_sample":
        base["pass"] = "one RFN.predict inner step: ConvLSTM cell + ListGlow.sample (T=0.7), autoregressive"
    if workload == "rfn_D_sample":
        base.update({"config": "D (main_rfn.py defaults)", "sequences_per_gpu": D["B"], "frames": "3x64x64, 2 context + 28 predicted",
                     "L": D["L"], "K": D["K"], "hidden": D["hidden"],
                     "pass": "28 autoregressive steps of ConvLSTM cell + ListGlow.sample (T=0.7)"})
    if workload == "glow_cfg1":
        base.update({"config": "Glow L=3 K=8 hidden 256, 1x32x32, batch 16, unconditional", "pass": "log_prob forward"})
    if workload == "convlstm_cfg2":
        base.update({"config": "ConvLSTM 64 -> 64, 3x3, 64x64 maps, T=10, batch 32", "pass": "forward"})
    return base


# ------------------------------------------------------------------------------------------------
# GPU arm helpers
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
            d = agg.setdefault(name, {"launches": 0, "ms": 0.0, "flops": 0.0, "flops_padded": 0.0, "bytes": 0.0})
            d["launches"] += 1
            d["ms"] += ms
            if meta:
                d["bytes"] += meta.get("bytes", 0.0)
            if meta and "flops" not in meta:   # bandwidth-bound kernel: algorithmic bytes only
                e = self.elementwise.setdefault(name, {})
                ee = e.setdefault(meta["bytes"], {"launches": 0, "ms": 0.0})
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
    for Cm in (4, 12, 48):
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


class Env:
    """Process / device context of the GPU arm."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL's INFO lines (communicator ranks, NVLS / ring choice) go to per-rank files next to the bench output; stdout
            # stays one JSON line
            if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
                os.environ["NCCL_DEBUG"] = "INFO"
                os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,GRAPH")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=self.dev)
        self.args = args
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        t = torch.tensor([ms], device=self.dev, dtype=torch.float32)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def time_steps(self, fn, steps, warmup=3, sampler=False):
        """W untimed warm-ups, then EXACTLY `steps` calls between barrier+synchronize, CUDA events, max over ranks (ms/step)."""
        for _ in range(max(warmup, 0)):
            fn()
        smp = ClockSampler(self.local) if (sampler and self.rank == 0) else None
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        self.barrier()
        t1 = time.time()
        ms = self.max_over_ranks(a.elapsed_time(b) / steps)
        return ms, (smp.summary(t0, t1) if smp else None)


def count_launches(rf, fn):
    """Launches of OUR kernels in one eager call of fn (graph replays do not pass through the ctypes counter)."""
    l0 = rf._lib.launches
    fn()
    return rf._lib.launches - l0


class HostFeeder:
    """Double-buffered host -> device input feed for a graph that reads STATIC device tensors: the H2D copy of step i+1
    runs on a copy stream into staging buffers while step i computes; a device-to-device copy at the head of step i+1
    moves them into the static inputs (what a pinned-memory data loader with a prefetch stream does)."""

    def __init__(self, host_tensors, static_tensors):
        self.host = [t.pin_memory() for t in host_tensors]
        self.static = static_tensors
        self.staging = [torch.empty_like(s) for s in static_tensors]
        self.copy_stream = torch.cuda.Stream()
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host)
        self.free = None     # event: staging consumed by the last D2D

    def prefetch(self):
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if self.free is not None:
                self.copy_stream.wait_event(self.free)
            else:
                self.copy_stream.wait_stream(main)
            for d, h in zip(self.staging, self.host):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return ev

    def consume(self, ev):
        main = torch.cuda.current_stream()
        main.wait_event(ev)
        for s, d in zip(self.static, self.staging):
            s.copy_(d, non_blocking=True)
        self.free = torch.cuda.Event()
        self.free.record(main)


def run_e2e(env, feeder, step_fn, result_fn, steps):
    """E2E loop: every step's inputs come from pinned host memory (prefetched one step ahead), result read back to the host."""
    def loop(n):
        ev = feeder.prefetch()
        for i in range(n):
            feeder.consume(ev)
            if i + 1 < n:
                ev = feeder.prefetch()
            step_fn()
            result_fn()
    loop(2)
    env.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loop(steps)
    b.record()
    env.barrier()
    return env.max_over_ranks(a.elapsed_time(b) / steps)


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def parity_block(rf, cfg, flow, lstm, dev, n=8):
    """Max-norm-relative error of z, logdet (nll) and bits/dim of OUR forward vs the CPU oracle on a full-depth sub-batch
    of n frames with the bench's weights, plus the ConvLSTM hidden state."""
    import oracle as O
    sd = {k: v.detach().cpu().clone() for k, v in flow.state_dict().items()}
    x, conds, z_part, feats = synth_inputs(cfg, n, n, 7)
    base = torch.cat([torch.randn(n, cfg["lstm_hidden"], 2, 2, generator=torch.Generator().manual_seed(3)), z_part], 1)
    noise = torch.rand(n, cfg["C"], 64, 64, generator=torch.Generator().manual_seed(4)) / 2 ** cfg["n_bits"]
    was_training = flow.training
    flow.eval()
    with torch.no_grad():
        z_ref, nll_ref = O.listglow_log_prob(x, conds, base, sd, cfg["L"], cfg["K"], cfg["n_bits"], noise=noise, learn_prior=True)
        z, nll = flow.log_prob(x.to(dev), [c.to(dev) for c in conds], base.to(dev), noise=noise.to(dev))
        conv = lstm.LSTMlayer.conv[0]
        h_ref = O.convlstm(feats[:, :3], conv.weight.detach().cpu(), conv.bias.detach().cpu())[0]
        h = lstm(feats[:, :3].to(dev))[0]
    flow.train(was_training)
    chw = cfg["C"] * 64 * 64
    bpd, bpd_ref = nll.cpu() / (math.log(2) * chw), nll_ref / (math.log(2) * chw)
    rel = lambda a, b: float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))   # noqa: E731
    return {"vs": "CPU oracle (oracle/, pinned to the reference by tests/golden)", "frames": n, "depth": f"L{cfg['L']} K{cfg['K']}",
            "z_max_norm_rel_err": rel(z, z_ref), "nll_max_norm_rel_err": rel(nll, nll_ref),
            "bits_per_dim_max_abs_err": float((bpd - bpd_ref).abs().max()), "bits_per_dim_ref_mean": float(bpd_ref.mean()),
            "convlstm_h_max_norm_rel_err": rel(h, h_ref), "gate": "1e-2 (bf16 convolutions, fp32 accumulate)"}


def roofline_block(env, rf, eager_step, what):
    """Per-kernel CUDA-event pass over one eager step (outside the timed regions): the dominant kernel family, its top
    (entry point, GEMM shape), and the fraction of the measured peaks."""
    peaks = env.peaks
    eager_step()
    torch.cuda.synchronize()
    kt = KernelTimer()
    rf._lib.tracer = kt
    torch.cuda.profiler.start()   # `ncu --profile-from-start off` captures exactly this one step
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    eager_step()
    ev1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    rf._lib.tracer = None
    table = kt.table()
    total_ms = sum(d["ms"] for d in table.values())
    burst = peaks.get("bf16_tflops", 1590.0)
    sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    fam_name, fam = max(table.items(), key=lambda kv: kv[1]["ms"])
    by_shape = []
    for (name, M, N, K), v in sorted(kt.shapes.items(), key=lambda kv: -kv[1]["ms"])[:8]:
        sec = v["ms"] / 1e3
        by_shape.append({"kernel": name, "M": M, "N": N, "K": K, "launches": v["launches"],
                         "avg_us": round(1e3 * v["ms"] / v["launches"], 1), "share_of_step": round(v["ms"] / total_ms, 4),
                         "tflops": round(v["flops"] / sec / 1e12, 1), "tensor_frac_burst": round(v["flops"] / sec / 1e12 / burst, 3),
                         "tensor_frac_sustained": round(v["flops"] / sec / 1e12 / sustained, 3),
                         "algorithmic_gbs": round(v["bytes"] / sec / 1e9, 1), "hbm_frac": round(v["bytes"] / sec / 1e9 / hbm_peak, 3)})
    # dominant kernel: the (entry point, shape) with the largest total time inside the dominant FAMILY (entry point)
    in_fam = [(k, v) for k, v in kt.shapes.items() if k[0] == fam_name]
    roofline = {"family": {"kernel": fam_name, "launches_per_step": fam["launches"], "ms": round(fam["ms"], 3),
                           "share_of_step": round(fam["ms"] / total_ms, 4),
                           "tflops": round(fam["flops"] / (fam["ms"] / 1e3) / 1e12, 1) if fam["flops"] else None,
                           "tensor_frac_burst": round(fam["flops"] / (fam["ms"] / 1e3) / 1e12 / burst, 3) if fam["flops"] else None,
                           "algorithmic_gbs": round(fam["bytes"] / (fam["ms"] / 1e3) / 1e9, 1),
                           "hbm_frac": round(fam["bytes"] / (fam["ms"] / 1e3) / 1e9 / hbm_peak, 3)}}
    if in_fam:
        tk, tv = max(in_fam, key=lambda kv: (kv[1]["ms"], kv[0]))
        sec = tv["ms"] / 1e3
        tfl, gbs = tv["flops"] / sec / 1e12, tv["bytes"] / sec / 1e9
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            hit = tj.get("|".join(str(v) for v in tk))
            traffic = hit["traffic_bytes"] if hit else None
        except (OSError, ValueError):
            pass
        bound = "hbm" if gbs / hbm_peak >= tfl / burst else "tensor"
        roofline.update({
            "kernel": f"{tk[0]} M={tk[1]} N={tk[2]} K={tk[3]}", "bound": bound,
            "achieved": round(gbs, 1) if bound == "hbm" else round(tfl, 1),
            "peak": hbm_peak if bound == "hbm" else burst, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
            "frac": round(gbs / hbm_peak if bound == "hbm" else tfl / burst, 3), "traffic": traffic,
            "peak_source": ("MEASURED_PEAKS.json: hbm_gbs; bf16_tflops (burst: the kernel is event-timed alone in an otherwise idle "
                            "eager step at full clocks)" if peaks else "fallback 6.65 TB/s, 1.59 PFLOP/s (B200_PROFILING.md)"),
            "tensor_frac_burst": round(tfl / burst, 3), "tensor_frac_sustained": round(tfl / sustained, 3),
            "hbm_frac": round(gbs / hbm_peak, 3),
            "algorithmic_bytes_per_launch": tv["bytes"] / tv["launches"], "algorithmic_flops_per_launch": tv["flops"] / tv["launches"],
            "share_of_step": round(tv["ms"] / total_ms, 4), "launches_per_step": tv["launches"],
            "avg_launch_us": round(1e3 * tv["ms"] / tv["launches"], 1)})
    else:   # bandwidth-bound dominant kernel
        gbs = fam["bytes"] / (fam["ms"] / 1e3) / 1e9
        roofline.update({"kernel": fam_name, "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(gbs / hbm_peak, 3), "traffic": None})
    roofline["step"] = what
    roofline["by_shape"] = by_shape
    roofline["instrumented_step_ms"] = {"sum_of_own_kernels": round(total_ms, 3), "wall_on_stream": round(ev0.elapsed_time(ev1), 3)}
    roofline["note"] = ("per-launch figures from CUDA events on the launching stream in an instrumented, single-stream eager step "
                        "(the timed step additionally overlaps the weight gradients on a second stream); algorithmic "
                        "bytes = activations in + out + weights; flops = 2*M*N*K on real (unpadded) channels")
    elementwise = []
    for name, by_bytes in kt.elementwise.items():
        nbytes = max(by_bytes)
        v = by_bytes[nbytes]
        us = 1e3 * v["ms"] / v["launches"]
        elementwise.append({"kernel": name, "algorithmic_bytes_per_launch": nbytes, "launches": v["launches"],
                            "avg_us": round(us, 1), "gbs": round(nbytes / us / 1e3, 1), "hbm_frac": round(nbytes / us / 1e3 / hbm_peak, 3)})
    elementwise.sort(key=lambda e: -e["algorithmic_bytes_per_launch"])
    roofline["elementwise_in_workload_largest_shape"] = elementwise[:12]
    kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4),
                   "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["flops"] else None}
               for k, v in sorted(table.items(), key=lambda kv: -kv[1]["ms"])}
    return roofline, kernels


def wl_rfn_train(env, rf, args, smooth=False):
    """Training step of the hot path, configuration J (optionally with the smoothing ConvLSTM of BASELINE config 4)."""
    cfg, dev, world = J, env.dev, env.world
    B, T = cfg["B"], cfg["T"]
    n_frames = B * (T - 1)
    hcn = cfg["lstm_hidden"]
    flow, lstm = build_models(cfg, n_frames)
    flow, lstm = flow.to(dev).train(), lstm.to(dev).train()
    slstm = None
    if smooth:
        torch.manual_seed(1)
        slstm = rf.ConvLSTM(cfg["lstm_in"] + hcn, hcn, [3, 3]).to(dev).train()
    hx, hconds, hz, hfeats = synth_inputs(cfg, n_frames, B, 1 + env.rank, dense=smooth)
    host = [hx, hz, hfeats] + hconds
    static = [t.to(dev) for t in host]
    tx, tz, tfeats, tconds = static[0], static[1], static[2], static[3:]
    params = list(flow.parameters()) + list(lstm.parameters()) + (list(slstm.parameters()) if slstm else [])
    opt = rf.FlatAdam(params, lr=1e-4, world_size=world)
    opt.attach(flow)      # per-level gradient slices are all-reduced while the shallower levels are still in the backward sweep

    def loss_fn():
        # as in RFN.loss: the recurrence's hidden states condition the flow's prior, so the flow's gradient w.r.t. its
        # base condition is back-propagated through all 19 ConvLSTM steps (and the smoothing ConvLSTM when enabled)
        hs, _, _ = lstm(tfeats)
        hflat = hs.reshape(n_frames, hcn, 2, 2)
        if slstm is not None:
            # a_t runs backward in time over cat[h_t, x-features] (RFN/RFN_new.py:142-153); it reaches the loss through the
            # encoder (outside the hot path), stood in for by a residual connection into the base condition
            a_in = torch.cat([hs, tfeats], 2).flip(1)
            a_s, _, _ = slstm(a_in)
            hflat = hflat + 0.1 * a_s.flip(1).reshape(n_frames, hcn, 2, 2)
        tbase = torch.cat([hflat, tz], 1)
        _, nll = flow.log_prob(tx, tconds, tbase)
        return nll.mean() / (math.log(2.0) * cfg["C"] * 64 * 64)

    def eager_step():
        opt.zero_grad()
        loss = loss_fn()
        loss.backward()
        opt.step()
        return loss.detach()

    def eager_step_local():   # the instrumented (rank-0 only) pass: same kernels, no collective
        from recurrent_flows_msc_b200.Flow import training as T
        opt.overlap = False   # ... including the per-level all-reduces the backward sweep would start
        side, T.WGRAD_SIDE_STREAM = T.WGRAD_SIDE_STREAM, False   # one stream: per-kernel event times are not inflated by overlap
        try:
            opt.zero_grad()
            loss = loss_fn()
            loss.backward()
            with torch.no_grad():
                opt.gather_grads()
                opt.apply()
        finally:
            opt.overlap = True
            T.WGRAD_SIDE_STREAM = side
        return loss.detach()

    first_loss = float(eager_step())
    launches_per_step = count_launches(rf, eager_step)
    train_step = eager_step if args.no_graph else rf.GraphedTrainStep(loss_fn, opt, warmup=2)
    ms, clocks = env.time_steps(train_step, args.steps, max(args.warmup, 3), sampler=True)
    # ---- e2e: the same step fed from pinned host memory, loss read back ------------------------------------------
    feeder = HostFeeder(host, static)
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    last = {}

    def step_fn():
        last["loss"] = train_step()

    def result_fn():
        loss_host.copy_(last["loss"].reshape(1), non_blocking=True)
    ms_e2e = run_e2e(env, feeder, step_fn, result_fn, args.steps)
    torch.cuda.synchronize()
    last_loss = float(loss_host[0])
    assert math.isfinite(last_loss), "non-finite training loss"
    out = {"value": world * n_frames / (ms / 1e3), "ms_per_step": ms, "clocks": clocks,
           "e2e": {"value": world * n_frames / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": feeder.h2d_bytes,
                   "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e,
                   "how": "pinned host -> device copy of step i+1 on a copy stream overlaps step i; loss copied back every step"},
           "gpu_launches": launches_per_step * args.steps,
           "training": {"own_kernel_launches_per_step": launches_per_step, "parameters": opt.n,
                        "launch": "eager" if args.no_graph else train_step.mode,
                        "allreduce_bytes_per_step": opt.n_pad * 4 if world > 1 else 0,
                        "allreduce": "per-level slices of the flat gradient, asynchronous, overlapped with the backward sweep; "
                                     "level 1 + ConvLSTM at the end" if world > 1 else "none (one replica)",
                        "loss_first_step": first_loss, "loss_last_step": last_loss,
                        "peak_memory_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}}
    state = dict(flow=flow, lstm=lstm, opt=opt, eager_step=eager_step_local, static=static, host=host, n_frames=n_frames)
    return out, state


def wl_rfn_fwd(env, rf, args, flow, lstm, static, host):
    """Forward (density evaluation) of the same models: 19 ConvLSTM steps + time-batched ListGlow.log_prob, no grad."""
    cfg, dev, world = J, env.dev, env.world
    B, T = cfg["B"], cfg["T"]
    n_frames = B * (T - 1)
    hcn = cfg["lstm_hidden"]
    flow.eval(); lstm.eval()

    def hot_path(dx, dz, dfeats, *dconds):
        with torch.no_grad():
            hs, h_last, _ = lstm(dfeats)
            base = torch.cat([hs.reshape(n_frames, hcn, 2, 2), dz], 1)
            _, nll = flow.log_prob(dx, list(dconds), base)
        return nll, h_last

    for _ in range(3):
        nll, _ = hot_path(*static)
    torch.cuda.synchronize()
    assert torch.isfinite(nll).all(), "non-finite nll in warm-up"
    launches_per_step = count_launches(rf, lambda: hot_path(*static))
    if args.no_graph:
        step = lambda: hot_path(*static)   # noqa: E731
        gin = static
    else:
        graphed = rf.Graphed(hot_path, *static)
        gin = graphed.static_in
        step = lambda: graphed.graph.replay()   # noqa: E731
        graphed(*static)
    ms, clocks = env.time_steps(step, args.steps, 3, sampler=False)
    feeder = HostFeeder(host, list(gin))
    nll_host = torch.empty(n_frames, dtype=torch.float32).pin_memory()
    h_host = torch.empty(B, hcn, 2, 2, dtype=torch.float32).pin_memory()
    res = {}

    def step_fn():
        if args.no_graph:
            res["o"] = hot_path(*gin)
        else:
            graphed.graph.replay()
            res["o"] = graphed.out

    def result_fn():
        nll_host.copy_(res["o"][0], non_blocking=True)
        h_host.copy_(res["o"][1], non_blocking=True)
    ms_e2e = run_e2e(env, feeder, step_fn, result_fn, args.steps)
    return {"what": "density evaluation: 19 ConvLSTM steps + time-batched ListGlow.log_prob on 570 frames per GPU, no grad, CUDA-graph replay",
            "frames_per_s": world * n_frames / (ms / 1e3), "ms_per_step": ms, "own_kernel_launches_per_step": launches_per_step,
            "e2e_frames_per_s": world * n_frames / (ms_e2e / 1e3), "e2e_ms_per_step": ms_e2e,
            "h2d_bytes_per_step": feeder.h2d_bytes, "d2h_bytes_per_step": nll_host.numel() * 4 + h_host.numel() * 4}, hot_path


def wl_sample(env, rf, args, cfg, flow, lstm, n_pred, steps):
    """RFN.predict's autoregressive inner loop on the hot path: per predicted frame one ConvLSTM cell step and one
    ListGlow.sample (T=0.7) for B sequences; n_pred frames in sequence per step (state carried on the device)."""
    dev, world = env.dev, env.world
    B = cfg["B"]
    flow.eval(); lstm.eval()
    _, hconds, hz, hfeats = synth_inputs(cfg, B, B, 11 + env.rank)
    conds = [c.to(dev) for c in hconds]
    zt = hz.to(dev)
    feat = hfeats.to(dev)                       # [B,1,lstm_in,2,2]: extractor features of the previous prediction (f2, not timed)
    hcn = cfg["lstm_hidden"]
    state = [torch.zeros(B, hcn, 2, 2, device=dev), torch.zeros(B, hcn, 2, 2, device=dev)]

    def one_frame(c_list, z_t, f_t, h, c):
        with torch.no_grad():
            _, h2, c2 = lstm(f_t, h, c)
            xs = flow.sample(None, list(c_list), torch.cat([h2, z_t], 1), num_samples=B, temperature=0.7)
        return xs, h2, c2

    launches = count_launches(rf, lambda: one_frame(conds, zt, feat, *state))
    for _ in range(2):
        one_frame(conds, zt, feat, *state)
    if args.no_graph:
        def frame():
            xs, h2, c2 = one_frame(conds, zt, feat, *state)
            state[0], state[1] = h2, c2
            return xs
    else:
        graphed = rf.Graphed(one_frame, conds, zt, feat, state[0], state[1])

        def frame():
            graphed.graph.replay()
            xs, h2, c2 = graphed.out
            graphed.static_in[3].copy_(h2)      # recurrent state stays on the device
            graphed.static_in[4].copy_(c2)
            return xs

    def step():
        for _ in range(n_pred):
            xs = frame()
        return xs
    ms, _ = env.time_steps(step, steps, 3)
    # e2e: conditions / latent / features of every frame come from pinned host memory, every predicted frame goes back
    # to the host (RFN.predict keeps `predictions` on the CPU, RFN/RFN_new.py:265,356)
    host = [t.pin_memory() for t in [hz, hfeats] + hconds]
    out_host = torch.empty(B, cfg["C"], 64, 64).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host)

    def step_e2e():
        for _ in range(n_pred):
            if args.no_graph:
                zt.copy_(host[0], non_blocking=True); feat.copy_(host[1], non_blocking=True)
                for d_, h_ in zip(conds, host[2:]):
                    d_.copy_(h_, non_blocking=True)
            else:
                graphed.static_in[1].copy_(host[0], non_blocking=True)
                graphed.static_in[2].copy_(host[1], non_blocking=True)
                for d_, h_ in zip(graphed.static_in[0], host[2:]):
                    d_.copy_(h_, non_blocking=True)
            out_host.copy_(frame(), non_blocking=True)
    ms_e2e, _ = env.time_steps(step_e2e, steps, 2)
    assert torch.isfinite(frame()).all()
    return {"what": f"{n_pred} autoregressive frame(s) per step: ConvLSTM cell + ListGlow.sample (config {cfg['name']}, T=0.7) for {B} sequences per GPU",
            "frames_per_s": world * B * n_pred / (ms / 1e3), "ms_per_step": ms, "ms_per_frame": ms / n_pred,
            "own_kernel_launches_per_frame": launches,
            "e2e_frames_per_s": world * B * n_pred / (ms_e2e / 1e3), "e2e_ms_per_step": ms_e2e,
            "h2d_bytes_per_step": h2d * n_pred, "d2h_bytes_per_step": out_host.numel() * 4 * n_pred,
            "launch": "eager" if args.no_graph else "CUDA graph replay per frame"}


def emit(env, args, workload, res, extra):
    world = env.world
    out = {"metric": METRICS[workload], "value": res["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": workload_config(workload, world, args),
           "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": res.get("clocks")}
    out.update(extra)
    print(json.dumps(out))


def run_ours(args):
    import recurrent_flows_msc_b200 as rf
    env = Env(args)
    dev, world, rank = env.dev, env.world, env.rank
    wl = args.workload
    extra = {}
    if wl in ("rfn_J_train", "rfn_J_smooth_D3"):
        res, st = wl_rfn_train(env, rf, args, smooth=wl == "rfn_J_smooth_D3")
        extra["training"] = res.pop("training")
        roof = kernels = None
        if rank == 0:
            roof, kernels = roofline_block(env, rf, st["eager_step"], "training step (eager, instrumented)")
        if wl == "rfn_J_train" and not args.only_headline:
            # the same (trained-for-a-few-steps) models in the other two passes the metric names
            rf.invalidate_caches()
            fwd, hot_path = wl_rfn_fwd(env, rf, args, st["flow"], st["lstm"], st["static"], st["host"])
            extra["forward"] = fwd
            extra["sampling"] = wl_sample(env, rf, args, J, st["flow"], st["lstm"], 1, max(args.steps, 10))
            if rank == 0:
                extra["parity"] = parity_block(rf, J, st["flow"], st["lstm"], dev)
                froof, fkern = roofline_block(env, rf, lambda: hot_path(*st["static"]), "forward (density evaluation) step")
                froof.pop("elementwise_in_workload_largest_shape", None)
                extra["forward"]["roofline"] = froof
                extra["forward"]["kernels_ms_per_step"] = fkern
        if rank == 0:
            roof["elementwise_at_scale"] = elementwise_at_scale(dev, env.peaks.get("hbm_gbs", 6650.0))
            extra["roofline"] = roof
            extra["kernels_ms_per_step"] = kernels
    elif wl == "rfn_J_fwd":
        cfg = J
        n_frames = cfg["B"] * (cfg["T"] - 1)
        flow, lstm = build_models(cfg, n_frames)
        flow, lstm = flow.to(dev), lstm.to(dev)
        hx, hconds, hz, hfeats = synth_inputs(cfg, n_frames, cfg["B"], 1 + rank)
        host = [hx, hz, hfeats] + hconds
        static = [t.to(dev) for t in host]
        fwd, hot_path = wl_rfn_fwd(env, rf, args, flow, lstm, static, host)
        res = {"value": fwd["frames_per_s"], "ms_per_step": fwd["ms_per_step"], "gpu_launches": fwd["own_kernel_launches_per_step"] * args.steps,
               "e2e": {"value": fwd["e2e_frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": fwd["h2d_bytes_per_step"],
                       "d2h_bytes_per_step": fwd["d2h_bytes_per_step"]}}
        if rank == 0:
            extra["parity"] = parity_block(rf, J, flow, lstm, dev)
            extra["roofline"], extra["kernels_ms_per_step"] = roofline_block(env, rf, lambda: hot_path(*static), "forward step")
    elif wl in ("rfn_J_sample", "rfn_D_sample"):
        cfg = J if wl == "rfn_J_sample" else D
        n_pred = 1 if wl == "rfn_J_sample" else 28
        flow, lstm = build_models(cfg, cfg["B"])
        flow, lstm = flow.to(dev), lstm.to(dev)
        s = wl_sample(env, rf, args, cfg, flow, lstm, n_pred, args.steps)
        res = {"value": s["frames_per_s"], "ms_per_step": s["ms_per_step"],
               "gpu_launches": s["own_kernel_launches_per_frame"] * n_pred * args.steps,
               "e2e": {"value": s["e2e_frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": s["h2d_bytes_per_step"],
                       "d2h_bytes_per_step": s["d2h_bytes_per_step"]}}
        extra["sampling"] = s
        if rank == 0:
            extra["parity"] = parity_block(rf, cfg, flow, lstm, dev, n=4)
            _, hconds, hz, hfeats = synth_inputs(cfg, cfg["B"], cfg["B"], 11)
            cd, zd, fd = [c.to(dev) for c in hconds], hz.to(dev), hfeats.to(dev)

            def eager():
                with torch.no_grad():
                    _, h2, _ = lstm(fd)
                    flow.sample(None, cd, torch.cat([h2, zd], 1), num_samples=cfg["B"], temperature=0.7)
            extra["roofline"], extra["kernels_ms_per_step"] = roofline_block(env, rf, eager, "one predicted frame (eager)")
    elif wl == "glow_cfg1":
        flow, sd, x, conds = cfg1_model()
        flow = flow.to(dev)
        xd, cd = x.to(dev), [c.to(dev) for c in conds]

        def fwd(xx):
            with torch.no_grad():
                return flow.log_prob(xx, cd, None)
        launches = count_launches(rf, lambda: fwd(xd))
        g = rf.Graphed(fwd, xd)
        ms, clocks = env.time_steps(lambda: g.graph.replay(), args.steps, 3, sampler=True)
        hx = x.pin_memory()
        nll_host = torch.empty(16).pin_memory()

        def e2e():
            g.static_in[0].copy_(hx, non_blocking=True)
            g.graph.replay()
            nll_host.copy_(g.out[1], non_blocking=True)
        ms_e2e, _ = env.time_steps(e2e, args.steps, 3)
        res = {"value": world * 16 / (ms / 1e3), "ms_per_step": ms, "gpu_launches": launches * args.steps, "clocks": clocks,
               "e2e": {"value": world * 16 / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": hx.numel() * 4, "d2h_bytes_per_step": 64}}
        if rank == 0:
            import oracle as O
            noise = torch.rand(16, 1, 32, 32) / 256
            with torch.no_grad():
                z_ref, nll_ref = O.listglow_log_prob(x, conds, None, sd, 3, 8, 8, noise=noise, learn_prior=False, make_conditional=False)
                z, nll = flow.log_prob(xd, cd, None, noise=noise.to(dev))
            extra["parity"] = {"vs": "CPU oracle", "z_max_norm_rel_err": float((z.cpu() - z_ref).abs().max() / z_ref.abs().max()),
                               "bits_per_dim_max_abs_err": float(((nll.cpu() - nll_ref) / (math.log(2) * 1024)).abs().max())}
            extra["roofline"], extra["kernels_ms_per_step"] = roofline_block(env, rf, lambda: fwd(xd), "cfg1 forward")
    elif wl == "convlstm_cfg2":
        torch.manual_seed(0)
        lstm = rf.ConvLSTM(64, 64, [3, 3]).to(dev).eval()
        hx = torch.randn(32, 10, 64, 64, 64, generator=torch.Generator().manual_seed(1 + rank))
        xd = hx.to(dev)

        def fwd(xx):
            with torch.no_grad():
                return lstm(xx)
        launches = count_launches(rf, lambda: fwd(xd))
        g = rf.Graphed(fwd, xd)
        ms, clocks = env.time_steps(lambda: g.graph.replay(), args.steps, 3, sampler=True)
        hxp = hx.pin_memory()
        h_host = torch.empty(32, 64, 64, 64).pin_memory()

        def e2e():
            g.static_in[0].copy_(hxp, non_blocking=True)
            g.graph.replay()
            h_host.copy_(g.out[1], non_blocking=True)
        ms_e2e, _ = env.time_steps(e2e, args.steps, 2)
        res = {"value": world * 320 / (ms / 1e3), "ms_per_step": ms, "gpu_launches": launches * args.steps, "clocks": clocks,
               "e2e": {"value": world * 320 / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": hxp.numel() * 4,
                       "d2h_bytes_per_step": h_host.numel() * 4}}
        if rank == 0:
            import oracle as O
            conv = lstm.LSTMlayer.conv[0]
            with torch.no_grad():
                o_ref = O.convlstm(hx[:2, :4], conv.weight.detach().cpu(), conv.bias.detach().cpu())[0]
                o = lstm(xd[:2, :4].contiguous())[0]
            extra["parity"] = {"vs": "CPU oracle", "h_max_norm_rel_err": float((o.cpu() - o_ref).abs().max() / o_ref.abs().max())}
            extra["roofline"], extra["kernels_ms_per_step"] = roofline_block(env, rf, lambda: fwd(xd), "cfg2 forward (10 cell steps)")
    else:
        raise SystemExit(f"unknown workload {wl}")
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            ref_frames = {"rfn_J_train": 60, "rfn_J_smooth_D3": 60, "rfn_J_fwd": 120, "rfn_J_sample": 30, "rfn_D_sample": 8,
                          "glow_cfg1": 16, "convlstm_cfg2": 40}[wl]
            extra["cpu_baseline"] = cpu_baseline_subprocess(wl, ref_frames)
        else:
            extra["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": None, "kind": "skipped",
                                     "sample": "timed at N=1 only (see the --impl reference arm)"}
        emit(env, args, wl, res, extra)
    if world > 1:
        # every rank is done; leave without tearing the communicator down under live CUDA graphs that contain collectives
        # (destroy_process_group() has been seen to hang in that state)
        torch.cuda.synchronize()
        env.dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rfn_J_train", choices=WORKLOADS)
    ap.add_argument("--ref-frames", type=int, default=None,
                    help="reference arm: frames per step (default: a bounded sample, 150 of the 570 frames for rfn_J_*)")
    ap.add_argument("--only-headline", action="store_true", help="skip the forward / sampling / parity sections of rfn_J_train")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the in-run CPU baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every launch from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
