"""Drop-in check against the real reference (/root/reference in the build container, baseline/_ref on the GPU box).

Builds the reference's RFN (RFN/RFN_new.py, main_rfn.py defaults) twice -- stock, and after
recurrent_flows_msc_b200.install_into(Flow, Utils) -- and checks the two models expose identical
state_dict keys and shapes, i.e. a reference checkpoint loads into the B200-backed model."""
import importlib
import os
import re
import sys
from unittest.mock import MagicMock

import pytest
import torch

from ref_helpers import reference_dir

REF = reference_dir() or "/nonexistent"   # /root/reference here, the staged baseline/_ref on the GPU box
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "RFN")), reason="reference checkout not present")


def reference_args():
    src = open(os.path.join(REF, "main_rfn.py")).read()
    head = src[:src.index("if __name__")]
    body = src[src.index("if __name__"):]
    body = body[body.index("\n") + 1:body.index("args = parser.parse_args()")]
    body = "\n".join(l[4:] if l.startswith("    ") else l for l in body.split("\n"))
    ns = {}
    exec(re.sub(r"^from .*$|^import (?!argparse).*$", "", head, flags=re.M) + "\nimport argparse\n" + body, ns)
    args = ns["parser"].parse_args([])
    args.batch_size = 2
    args.x_dim = [2] + list(args.x_dim[1:])
    args.condition_dim = [2] + list(args.condition_dim[1:])
    return args


def test_install_into_keeps_rfn_state_dict():
    sys.dont_write_bytecode = True
    for m in ["matplotlib", "matplotlib.pyplot", "imageio", "torchfile", "parse"]:
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, REF)
    try:
        import Flow
        import Utils
        args = reference_args()
        rfn_mod = importlib.import_module("RFN.RFN_new")
        torch.manual_seed(0)
        stock = rfn_mod.RFN(args)
        ref_sd = {k: tuple(v.shape) for k, v in stock.state_dict().items()}
        import recurrent_flows_msc_b200 as rfk
        saved = {n: getattr(Flow, n) for n in ("ListGlow",)}, {n: getattr(Utils, n) for n in ("ConvLSTM", "ConvLSTMLayer")}
        patched = rfk.install_into(Flow, Utils)
        assert "Flow.ListGlow" in patched and "Utils.ConvLSTM" in patched
        try:
            rfn_mod = importlib.reload(rfn_mod)
            torch.manual_seed(0)
            ours = rfn_mod.RFN(args)
            assert isinstance(ours.flow, rfk.ListGlow) and isinstance(ours.lstm, rfk.ConvLSTM)
            our_sd = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
            assert our_sd == ref_sd
            ours.load_state_dict(stock.state_dict())   # a reference checkpoint loads unchanged
        finally:
            importlib.reload(importlib.import_module("Flow.glow_modules"))
            importlib.reload(importlib.import_module("Flow.glow"))
            importlib.reload(Flow)
            importlib.reload(importlib.import_module("Utils.modules"))
            importlib.reload(Utils)
    finally:
        sys.path.remove(REF)
        for name in [n for n in sys.modules if n.split(".")[0] in ("Flow", "Utils", "RFN")]:
            del sys.modules[name]


def test_reference_checkpoint_file_round_trip(tmp_path):
    """A checkpoint written the way the reference's Solver.checkpoint does (RFN/trainer.py:277-289: model_state_dict,
    optimizer_state_dict, counters and the pickled argparse Namespace in one torch.save) by a STOCK model loads into the
    B200-backed model built from the Namespace inside the file, as main_rfn.py:7-12 + Solver.load (:302-315) do on
    resume, and every tensor arrives bit for bit."""
    sys.dont_write_bytecode = True
    for m in ["matplotlib", "matplotlib.pyplot", "imageio", "torchfile", "parse"]:
        sys.modules.setdefault(m, MagicMock())
    sys.path.insert(0, REF)
    try:
        import Flow
        import Utils
        args = reference_args()
        rfn_mod = importlib.import_module("RFN.RFN_new")
        torch.manual_seed(3)
        stock = rfn_mod.RFN(args)
        with torch.no_grad():
            for p in stock.parameters():
                p.add_(torch.randn_like(p) * 0.01)
        opt = torch.optim.Adam(stock.parameters(), lr=args.learning_rate)
        path = str(tmp_path / "rfn.pt")
        torch.save({"epoch": 7, "model_state_dict": stock.state_dict(), "optimizer_state_dict": opt.state_dict(), "loss": 1.5,
                    "kl_loss": [0.1], "recon_loss": [1.4], "losses": [1.5], "plot_counter": 2, "annealing_counter": 123,
                    "bits_per_dim": [3.2], "args": args}, path)
        import recurrent_flows_msc_b200 as rfk
        rfk.install_into(Flow, Utils)
        try:
            rfn_mod = importlib.reload(rfn_mod)
            ckpt = torch.load(path, map_location="cpu", weights_only=False)
            ours = rfn_mod.RFN(ckpt["args"])                    # the resume path rebuilds the model from the stored args
            missing = ours.load_state_dict(ckpt["model_state_dict"])
            assert not missing.missing_keys and not missing.unexpected_keys
            for k, v in stock.state_dict().items():
                assert torch.equal(ours.state_dict()[k], v), k
            assert ckpt["epoch"] == 7 and ckpt["annealing_counter"] == 123
        finally:
            importlib.reload(importlib.import_module("Flow.glow_modules"))
            importlib.reload(importlib.import_module("Flow.glow"))
            importlib.reload(Flow)
            importlib.reload(importlib.import_module("Utils.modules"))
            importlib.reload(Utils)
    finally:
        sys.path.remove(REF)
        for name in [n for n in sys.modules if n.split(".")[0] in ("Flow", "Utils", "RFN")]:
            del sys.modules[name]
