"""RFN-level drop-in: the reference's own RFN (RFN/RFN_new.py, configuration J of RFN/default_rfn_job.sh) built twice --
stock, and with recurrent_flows_msc_b200.install_into(Flow, Utils) -- on identical weights, inputs and seeds.

* CPU (not gpu): the time-batched loss driver equals the unchanged RFN.loss on the stock model (driver logic only).
* GPU: loss() (kl, nll), its gradients, predict() and the time-batched driver of the patched model against the stock
  model running PyTorch's own CUDA kernels.

The reference comes from /root/reference (build container) or baseline/_ref (staged by __graft_entry__.build(); travels
to the GPU box).  Tolerances: the flow's convolutions run in bf16 with fp32 accumulation (gate 1e-2, BASELINE.json)."""
import importlib
import os
import sys

import pytest
import torch

from ref_helpers import job_script_args, purge_reference_modules, reference_dir, stub_optional_imports

REF = reference_dir()
pytestmark = pytest.mark.skipif(REF is None, reason="reference checkout / baseline/_ref not present")


def _import_reference():
    sys.dont_write_bytecode = True
    stub_optional_imports()
    purge_reference_modules()
    sys.path.insert(0, REF)
    import Flow
    import Utils
    rfn_mod = importlib.import_module("RFN.RFN_new")
    return Flow, Utils, rfn_mod


def _cleanup():
    if REF in sys.path:
        sys.path.remove(REF)
    purge_reference_modules()


def _data(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(B, T, 1, 64, 64, generator=g) * (torch.rand(B, T, 1, 64, 64, generator=g) < 0.3).float()
    return torch.floor(u * 256) / 256 - 0.5


def test_time_batched_loss_equals_rfn_loss_cpu(monkeypatch):
    """Driver logic on the stock model: with the dequantisation draw removed (its ORDER is the one thing time-batching
    changes) one log_prob call on B*(T-1) frames gives the same (kl_free_bits, kl, nll) and the same gradients."""
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self, raising=False)
    Flow, Utils, rfn_mod = _import_reference()
    try:
        from recurrent_flows_msc_b200.rfn_driver import time_batched_loss
        B, T = 2, 4
        args = job_script_args(REF, B, ["--K", "1", "--n_units_affine", "16", "--n_units_prior", "16"])
        torch.manual_seed(0)
        rfn = rfn_mod.RFN(args).train()
        with torch.no_grad():
            for n, p in rfn.named_parameters():
                if "flow" in n:
                    p.add_(torch.randn_like(p) * 0.02)
        for m in rfn.flow.modules():            # skip the data-dependent ActNorm init: it would see B vs B*(T-1) frames
            if hasattr(m, "initialized"):
                m.initialized.fill_(1)
        flow = rfn.flow

        def no_noise(x):
            b, c, h, w = x.shape
            return x, -float(torch.log(torch.tensor(2.0 ** flow.n_bits))) * c * h * w * torch.ones(b)
        monkeypatch.setattr(flow, "uniform_binning_correction", no_noise)
        x = _data(B, T)
        torch.manual_seed(1)
        ref = rfn.loss(x, 0)
        (ref[0] + ref[2]).backward()
        g_ref = {n: p.grad.clone() for n, p in rfn.named_parameters() if p.grad is not None}
        rfn.zero_grad()
        torch.manual_seed(1)
        got = time_batched_loss(rfn, x, 0)
        (got[0] + got[2]).backward()
        for a, b in zip(ref, got):
            torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)
        assert "log_prob" not in flow.__dict__          # the method is restored
        for n, p in rfn.named_parameters():
            if n in g_ref:
                torch.testing.assert_close(p.grad, g_ref[n], rtol=2e-3, atol=1e-5, msg=lambda m, n=n: f"{n}: {m}")
    finally:
        _cleanup()


def _max_rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.mark.gpu
def test_rfn_loss_and_predict_stock_vs_patched():
    """INTEGRATION.md sections 2/3b executed on hardware: RFN.loss is 19 log_prob calls (T=20) at batch B in train mode
    with a python-int logdet and batch-baked initial states; predict(10, 10) is the autoregressive sampling path."""
    import recurrent_flows_msc_b200 as rfk
    Flow, Utils, rfn_mod = _import_reference()
    try:
        B, T = 4, 20
        args = job_script_args(REF, B)
        torch.manual_seed(0)
        stock = rfn_mod.RFN(args).cuda().train()
        with torch.no_grad():
            g = torch.Generator().manual_seed(5)
            # trained-like: zero-init Conv2dZeros / realnvp scale would make the flow trivial.  The scale keeps the 50-step
            # reverse flow in a trained model's regime (samples within the image range; 2.5x larger perturbations give
            # |x| ~ 1e5, where the comparison measures the conditioning of the random map: tools/dropin_debug.py)
            for n, p in stock.named_parameters():
                if n.startswith("flow."):
                    p.add_((torch.randn(p.shape, generator=g) * (0.004 if "conv.weight" in n else 0.02)).cuda())
        sd0 = {k: v.clone() for k, v in stock.state_dict().items()}
        x = _data(B, T).cuda()

        torch.manual_seed(11)
        kl_fb_s, kl_s, nll_s = stock.loss(x, 0)          # first training call: data-dependent ActNorm init
        (kl_fb_s + nll_s).backward()
        sd1 = {k: v.clone() for k, v in stock.state_dict().items()}     # with initialised ActNorms

        patched_names = rfk.install_into(Flow, Utils)
        assert "Flow.ListGlow" in patched_names and "Utils.ConvLSTM" in patched_names
        rfn_mod2 = importlib.reload(rfn_mod)
        torch.manual_seed(0)
        ours = rfn_mod2.RFN(args).cuda().train()
        assert isinstance(ours.flow, rfk.ListGlow) and isinstance(ours.lstm, rfk.ConvLSTM)
        ours.load_state_dict(sd0)
        torch.manual_seed(11)
        kl_fb_o, kl_o, nll_o = ours.loss(x, 0)
        (kl_fb_o + nll_o).backward()
        chw_t = 64 * 64 * (T - 1)
        bpd_s, bpd_o = float(nll_s.detach()) / (0.6931 * chw_t), float(nll_o.detach()) / (0.6931 * chw_t)
        print(f"RFN.loss stock vs patched: nll {float(nll_s):.4f} / {float(nll_o):.4f}  bits/dim {bpd_s:.5f} / {bpd_o:.5f}  "
              f"kl {float(kl_s):.5f} / {float(kl_o):.5f}")
        assert abs(float(nll_o) - float(nll_s)) <= 1e-2 * abs(float(nll_s)) + 1e-3 * chw_t * 0.6931   # 1e-2 rel or 1e-3 bits/dim
        assert abs(float(kl_o) - float(kl_s)) <= 2e-2 * abs(float(kl_s)) + 1e-3
        # data-dependent ActNorm init gave the same parameters
        for k in ("flow.glow_frame.1.norm.logs", "flow.glow_frame.1.affine.net.0.norm_type.logs", "flow.prior.0.norm_type.bias"):
            assert _max_rel(ours.state_dict()[k], sd1[k]) < 2e-2, k
        # gradients reach the torch modules around the hot path (upscaler, extractor, prior net) through OUR backward
        gs = dict(stock.named_parameters())
        checked = 0
        for n, p in ours.named_parameters():
            if p.grad is None or gs[n].grad is None or float(gs[n].grad.abs().max()) < 1e-8:
                continue
            if n.startswith(("upscaler.", "extractor.", "lstm.", "flow.prior.", "flow.glow_frame.1.", "flow.glow_frame.45.")):
                c = _cos(p.grad, gs[n].grad)
                assert c > 0.9, f"gradient of {n}: cosine {c:.4f} vs stock autograd"
                checked += 1
        assert checked > 20

        # time-batched driver on the patched model (same weights, same seed): same loss up to the dequantisation order
        from recurrent_flows_msc_b200.rfn_driver import time_batched_loss
        ours.load_state_dict(sd1)
        ours.zero_grad()
        torch.manual_seed(12)
        _, kl_t, nll_t = time_batched_loss(ours, x, 0)
        stock.zero_grad()
        torch.manual_seed(12)
        _, kl_s2, nll_s2 = stock.loss(x, 0)
        print(f"time-batched: nll {float(nll_t):.4f} vs stock {float(nll_s2):.4f}")
        assert abs(float(nll_t) - float(nll_s2)) <= 2e-2 * abs(float(nll_s2)) + 2e-3 * chw_t * 0.6931

        # predict: identical weights (incl. the initialised ActNorms and BatchNorm running statistics), eval mode
        stock.load_state_dict(sd1); ours.load_state_dict(sd1)
        stock.eval(); ours.eval()
        with torch.no_grad():
            torch.manual_seed(21)
            true_s, pred_s = stock.predict(x, 10, 10)
            torch.manual_seed(21)
            true_o, pred_o = ours.predict(x, 10, 10)
        assert pred_o.shape == pred_s.shape == (10, B, 1, 64, 64) and torch.isfinite(pred_o).all()
        assert torch.equal(true_o, true_s)
        e0 = _max_rel(pred_o[0], pred_s[0])
        e_all = float((pred_o - pred_s).abs().mean() / pred_s.abs().mean().clamp_min(1e-12))
        print(f"predict(10,10): first frame max-norm rel err {e0:.3e}, all frames mean abs rel {e_all:.3e}")
        assert e0 < 2e-2          # one pass through the reverse flow (bf16 gate 1e-2, doubled for the 50-step inverse)
        assert e_all < 5e-2       # ten autoregressive passes: errors feed back through extractor and ConvLSTM
    finally:
        for name in ("Flow.glow_modules", "Flow.glow", "Flow", "Utils.modules", "Utils"):
            if name in sys.modules:
                importlib.reload(sys.modules[name])
        _cleanup()
