"""the reference's feature extractor (VGG_downscaler), condition upscaler (VGG_upscaler) and Gaussian parameter
nets (SimpleParamNet) -- Utils/modules.py:43-244 -- with recurrent_flows_msc_b200.accelerate_scalers(): eval-mode outputs
against the same modules' own PyTorch forward (bf16 convolutions, gate 1e-2 of the max-norm), training mode untouched."""
import importlib
import sys

import pytest
import torch

from ref_helpers import job_script_args, purge_reference_modules, reference_dir, stub_optional_imports

REF = reference_dir()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(REF is None, reason="reference checkout / baseline/_ref not present")]


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_scalers_eval_forward_matches_reference_modules():
    import recurrent_flows_msc_b200 as rfk
    sys.dont_write_bytecode = True
    stub_optional_imports()
    purge_reference_modules()
    sys.path.insert(0, REF)
    try:
        rfn_mod = importlib.import_module("RFN.RFN_new")
        B = 6
        args = job_script_args(REF, B, ["--K", "1"])
        torch.manual_seed(0)
        rfn = rfn_mod.RFN(args).cuda()
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for name, buf in rfn.named_buffers():                      # non-trivial BatchNorm statistics
                if name.endswith("running_var"):
                    buf.copy_((torch.rand(buf.shape, generator=g) + 0.5).cuda())
                elif name.endswith("running_mean"):
                    buf.copy_((torch.randn(buf.shape, generator=g) * 0.2).cuda())
        rfn.eval()
        x = (torch.rand(B, 1, 64, 64, generator=g) - 0.5).cuda()
        hz = torch.randn(B, 256, 2, 2, generator=g).cuda()
        with torch.no_grad():
            feats_ref = rfn.extractor(x)
            ups_ref = rfn.upscaler(hz, skip_list=list(feats_ref))
            prior_ref = rfn.prior(hz)
            enc_ref = rfn.encoder(torch.cat([hz, feats_ref[-1]], 1))
            patched = rfk.accelerate_scalers(rfn)
            assert {"extractor", "upscaler", "prior", "encoder"} <= set(patched)
            l0 = rfk._lib.launches
            feats = rfn.extractor(x)
            ups = rfn.upscaler(hz, skip_list=list(feats_ref))
            prior = rfn.prior(hz)
            enc = rfn.encoder(torch.cat([hz, feats_ref[-1]], 1))
            assert rfk._lib.launches - l0 >= 25                       # the convolutions went through librfk
        errs = [_rel(a, b) for a, b in zip(feats, feats_ref)] + [_rel(a, b) for a, b in zip(ups, ups_ref)]
        errs += [_rel(prior[0], prior_ref[0]), _rel(prior[1], prior_ref[1]), _rel(enc[0], enc_ref[0]), _rel(enc[1], enc_ref[1])]
        print("scalers: max-norm rel errors", ["%.2e" % e for e in errs])
        assert max(errs) < 1e-2
        # training mode / autograd: the module's own forward (identical tensors, gradients flow)
        rfn.train()
        xg = x.clone().requires_grad_()
        out = rfn.extractor(xg)
        out[-1].sum().backward()
        assert xg.grad is not None
        assert rfk.accelerate_scalers(rfn) == []                      # idempotent
    finally:
        if REF in sys.path:
            sys.path.remove(REF)
        purge_reference_modules()
