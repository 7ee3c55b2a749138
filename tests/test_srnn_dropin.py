"""SRNN-level drop-in of the ConvLSTM (BASELINE north_star: "drops into RFN.py, SRNN.py and glow.py"): the reference's own
SRNN (SRNN/SRNN.py, main_srnn.py defaults: two ConvLSTMs 256 -> 60 and 316 -> 60 on 8x8 maps, smoothing enabled) built
stock and with recurrent_flows_msc_b200.install_into(None, Utils), same weights, inputs and seeds: loss (kl, nll), the
gradients that flow back through OUR BPTT into the reference's conv encoders, and predict()."""
import importlib
import os
import sys

import pytest
import torch

from ref_helpers import purge_reference_modules, reference_args, reference_dir, stub_optional_imports

REF = reference_dir()
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(REF is None or not os.path.isfile(os.path.join(REF or "", "SRNN", "SRNN.py")),
                                 reason="reference checkout / baseline/_ref (with SRNN) not present")]


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def test_srnn_loss_gradients_predict_stock_vs_patched():
    import recurrent_flows_msc_b200 as rfk
    sys.dont_write_bytecode = True
    stub_optional_imports()
    purge_reference_modules()
    sys.path.insert(0, REF)
    try:
        import Utils
        srnn_mod = importlib.import_module("SRNN.SRNN")
        B, T = 6, 8
        args = reference_args(REF, ["--batch_size", str(B), "--x_dim", str(B), "1", "64", "64", "--condition_dim", str(B), "1", "64",
                                    "64", "--enable_smoothing"], script="main_srnn.py")
        torch.manual_seed(0)
        stock = srnn_mod.SRNN(args).cuda().train()
        sd0 = {k: v.clone() for k, v in stock.state_dict().items()}
        g = torch.Generator().manual_seed(1)
        x = (torch.floor(torch.rand(B, T, 1, 64, 64, generator=g) * 256) / 256 - 0.5).cuda()
        torch.manual_seed(5)
        kl_s, nll_s = stock.loss(x)
        (kl_s + nll_s).backward()

        assert "Utils.ConvLSTM" in rfk.install_into(None, Utils)
        srnn_mod2 = importlib.reload(srnn_mod)
        torch.manual_seed(0)
        ours = srnn_mod2.SRNN(args).cuda().train()
        assert isinstance(ours.lstm_h, rfk.ConvLSTM) and isinstance(ours.lstm_a, rfk.ConvLSTM)
        ours.load_state_dict(sd0)
        torch.manual_seed(5)
        kl_o, nll_o = ours.loss(x)
        (kl_o + nll_o).backward()
        print(f"SRNN.loss stock vs patched: kl {float(kl_s):.5f} / {float(kl_o):.5f}  nll {float(nll_s):.3f} / {float(nll_o):.3f}")
        assert abs(float(nll_o) - float(nll_s)) <= 1e-2 * abs(float(nll_s)) + 1e-2
        assert abs(float(kl_o) - float(kl_s)) <= 2e-2 * abs(float(kl_s)) + 1e-3
        gs = dict(stock.named_parameters())
        checked = 0
        gmax = max(float(q.grad.abs().max()) for q in gs.values() if q.grad is not None)
        for n, p in ours.named_parameters():
            # (conv biases in front of a BatchNorm have a mathematically zero gradient: rounding noise in both models)
            if p.grad is None or gs[n].grad is None or float(gs[n].grad.abs().max()) < 1e-5 * gmax:
                continue
            if n.startswith(("lstm_h.", "lstm_a.")) or (n.startswith("phi_x_t.") and n.endswith("weight")):
                c = _cos(p.grad, gs[n].grad)
                assert c > 0.98, f"gradient of {n}: cosine {c:.4f} vs stock autograd"
                checked += 1
        assert checked >= 6
        stock.eval(); ours.eval()
        ours.load_state_dict(stock.state_dict(), strict=False)   # (peephole tensors, if the reference registered any, are zeros)
        with torch.no_grad():
            torch.manual_seed(9)
            _, pred_s = stock.predict(x, 3, 5)
            torch.manual_seed(9)
            _, pred_o = ours.predict(x, 3, 5)
        # the decoder draws from a discretised mixture of logistics: a 1e-3 difference in h can flip a component choice, so
        # single pixels may differ by O(1) while the frames agree -- the criterion is the mean absolute difference
        err = float((pred_o - pred_s).abs().max() / pred_s.abs().max().clamp_min(1e-12))
        mean_err = float((pred_o - pred_s).abs().mean())
        frac = float(((pred_o - pred_s).abs() > 0.05).float().mean())
        print(f"SRNN.predict(3, 5): max-norm rel err {err:.3e}, mean abs diff {mean_err:.3e}, pixels off by > 0.05: {100 * frac:.2f} %")
        assert torch.isfinite(pred_o).all() and pred_o.shape == pred_s.shape and mean_err < 2e-2
    finally:
        for name in ("Utils.modules", "Utils"):
            if name in sys.modules:
                importlib.reload(sys.modules[name])
        if REF in sys.path:
            sys.path.remove(REF)
        purge_reference_modules()
