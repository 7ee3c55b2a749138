"""Pin oracle/ against fixtures generated from the reference's own modules
(tests/golden/make_golden.py).  CPU only."""
import math

import pytest
import torch

import oracle as O
from conftest import load_golden

TOL = dict(rtol=1e-5, atol=1e-5)


def close(a, b, **kw):
    tol = dict(TOL)
    tol.update(kw)
    torch.testing.assert_close(a, b, **tol)


def test_squeeze_bit_exact():
    g = load_golden("squeeze")
    assert torch.equal(O.squeeze2d(g["x"], False), g["y"])
    assert torch.equal(O.squeeze2d(g["y"], True), g["undo"])
    assert torch.equal(g["undo"], g["x"])
    # index law of SURVEY 8(a4)
    x = g["x"]
    y = O.squeeze2d(x)
    for c in range(x.shape[1]):
        for dy in range(2):
            for dx in range(2):
                assert torch.equal(y[:, 4 * c + 2 * dy + dx], x[:, c, dy::2, dx::2])


def test_actnorm_init_and_directions():
    g = load_golden("actnorm")
    bias, logs = O.actnorm_init(g["x"])
    close(bias, g["sd_after_init"]["bias"])
    close(logs, g["sd_after_init"]["logs"])
    y, ld = O.actnorm(g["x"], bias, logs, torch.zeros(3), False)
    close(y, g["y_init_train"])
    close(ld, g["logdet_init_train"])
    xr, ldr = O.actnorm(y, bias, logs, ld, True)
    close(xr, g["x_rev"])
    close(ldr, g["logdet_rev"], atol=1e-4)
    # eval-first-call quirk: parameters stay zero -> identity, flag set
    assert int(g["sd_eval_first"]["initialized"]) == 1
    assert torch.equal(g["sd_eval_first"]["logs"], torch.zeros(1, 5, 1, 1))
    y2, _ = O.actnorm(g["x"], torch.zeros(5), torch.zeros(5), None, False)
    close(y2, g["y_eval_first"])
    y3, ld3 = O.actnorm(g["x3"], bias, logs, 0.0, False)
    close(y3, g["y3"])
    close(ld3, g["logdet3"])
    assert ld3.dim() == 0


@pytest.mark.parametrize("name", ["invconv_lu", "invconv_plain"])
def test_invconv(name):
    g = load_golden(name)
    z, ld = O.invconv(g["x"], g["sd"], "", torch.zeros(2), False)
    close(z, g["z"])
    close(ld, g["logdet"])
    xr, ldr = O.invconv(z, g["sd"], "", ld, True)
    close(xr, g["x_rev"], atol=1e-4)
    close(ldr, g["logdet_rev"], atol=1e-4)


@pytest.mark.parametrize("clamp", ["realnvp", "glow", "softclamp", "none"])
def test_coupling(clamp):
    g = load_golden(f"coupling_{clamp}")
    z, ld = O.affine_coupling(g["x"], g["cond"], g["sd"], "", torch.zeros(2), False, g["clamp_type"], g["non_lin"])
    close(z, g["z"])
    close(ld, g["logdet"], atol=1e-4)
    xr, ldr = O.affine_coupling(z, g["cond"], g["sd"], "", ld, True, g["clamp_type"], g["non_lin"])
    close(xr, g["x_rev"], atol=1e-4)
    close(ldr, g["logdet_rev"], atol=1e-4)


@pytest.mark.parametrize("name", ["split2d_cond_softplus", "split2d_uncond_exp"])
def test_split2d(name):
    g = load_golden(name)
    z1, ld = O.split2d(g["x"], g["cond"], g["sd"], "", torch.zeros(2), False, None,
                       g["make_conditional"], g["clamp_function"])
    assert torch.equal(z1, g["z1"])
    close(ld, g["logdet"], atol=1e-4)
    xr, _ = O.split2d(z1, g["cond"], g["sd"], "", None, True, g["temperature"],
                      g["make_conditional"], g["clamp_function"], g["eps"])
    close(xr, g["x_rev"])


def test_glowstep():
    g = load_golden("glowstep")
    z, ld = O.glow_step(g["x"], g["cond"], g["sd"], "", torch.zeros(2), False)
    close(z, g["z"])
    close(ld, g["logdet"], atol=1e-4)
    xr, ldr = O.glow_step(z, g["cond"], g["sd"], "", ld, True)
    close(xr, g["x_rev"], atol=1e-4)
    close(ldr, g["logdet_rev"], atol=1e-4)
    close(xr, g["x"], atol=1e-4)  # bijection


def _kw(a):
    return dict(clamp_type=a["clamp_type"], non_lin=a["non_lin_glow"],
                make_conditional=a["make_conditional"], split2d_act=a["split2d_act"])


def test_listglow_cond():
    g = load_golden("listglow_cond")
    a = g["args"]
    z, ld = O.listglow_f(g["x"], g["cond"], g["sd"], a["L"], a["K"], 0.0, **_kw(a))
    close(z, g["z_f"], atol=1e-4)
    close(ld, g["logdet_f"], atol=1e-3)
    z2, nll = O.listglow_log_prob(g["x"], g["cond"], g["base"], g["sd"], a["L"], a["K"], a["n_bits"],
                                  noise=g["noise"], learn_prior=True, **_kw(a))
    close(z2, g["z_logprob"], atol=1e-4)
    close(nll, g["nll"], rtol=1e-5, atol=1e-3)
    close(O.bits_per_dim(nll, 256), g["bpd"], atol=1e-5)
    xs = O.listglow_sample(g["cond"], g["base"], g["sd"], a["L"], a["K"], g["eps_prior"], g["eps_split"],
                           g["temperature"], learn_prior=True, **_kw(a))
    close(xs, g["x_sample"], atol=1e-4)
    # g(f(x)) == x when the Split2d draw reproduces the z2 that was split off is covered by test_glowstep;
    # here check invertibility of the non-split part through the last level only
    assert math.isfinite(float(nll.sum()))


def test_listglow_uncond():
    g = load_golden("listglow_uncond")
    a = g["args"]
    conds = [torch.zeros(*s) for s in g["cond_sizes"]]
    z, nll = O.listglow_log_prob(g["x"], conds, None, g["sd"], a["L"], a["K"], a["n_bits"],
                                 noise=g["noise"], learn_prior=False, **_kw(a))
    close(z, g["z_logprob"], atol=1e-4)
    close(nll, g["nll"], rtol=1e-5, atol=1e-3)
    xs = O.listglow_sample(conds, None, g["sd"], a["L"], a["K"], g["eps_prior"], g["eps_split"],
                           g["temperature"], learn_prior=False, **_kw(a))
    close(xs, g["x_sample"], atol=1e-4)


def test_convlstm():
    g = load_golden("convlstm")
    w, b = g["sd"]["LSTMlayer.conv.0.weight"], g["sd"]["LSTMlayer.conv.0.bias"]
    out, h, c = O.convlstm(g["x"], w, b)
    close(out, g["out"])
    close(h, g["h"])
    close(c, g["c"])
    out2, h2, c2 = O.convlstm(g["x"][:, :1], w, b, g["h0"], g["c0"])
    close(out2, g["out2"])
    close(h2, g["h2"])
    close(c2, g["c2"])


def test_batchnormflow():
    g = load_golden("batchnormflow")
    lg, beta = g["sd0"]["log_gamma"], g["sd0"]["beta"]
    y, ld, rm, rv = O.batchnorm_flow(g["x"], lg, beta, torch.zeros_like(lg), torch.ones_like(lg), torch.zeros(4), False,
                                     training=True, momentum=g["momentum"])
    close(y, g["y_train"], atol=1e-4)
    close(ld, g["logdet_train"], atol=1e-4)
    close(rm, g["sd_after"]["running_mean"])
    close(rv, g["sd_after"]["running_var"])
    y2, ld2, _, _ = O.batchnorm_flow(g["x2"], lg, beta, rm, rv, torch.zeros(2), False)
    close(y2, g["y_eval"], atol=1e-4)
    close(ld2, g["logdet_eval"], atol=1e-4)
    xr, ldr, _, _ = O.batchnorm_flow(y2, lg, beta, rm, rv, ld2, True)
    close(xr, g["x_rev"], atol=1e-4)
    close(ldr, g["logdet_rev"], atol=1e-4)


def test_listglow_batchnorm():
    g = load_golden("listglow_batchnorm")
    a = g["args"]
    z, nll = O.listglow_log_prob(g["x"], g["cond"], g["base"], g["sd"], a["L"], a["K"], a["n_bits"],
                                 noise=g["noise"], learn_prior=True, **_kw(a))
    close(z, g["z_logprob"], atol=1e-4)
    close(nll, g["nll"], rtol=1e-5, atol=1e-3)
    xs = O.listglow_sample(g["cond"], g["base"], g["sd"], a["L"], a["K"], g["eps_prior"], g["eps_split"],
                           g["temperature"], learn_prior=True, **_kw(a))
    close(xs, g["x_sample"], atol=1e-4)


def test_listglow_batchnorm_training_mode_and_grads():
    """flow_norm='batchnorm' + base_norm='batchnorm' in train() mode: batch statistics, and autograd through the oracle
    gives the reference's gradients."""
    g = load_golden("listglow_batchnorm_train")
    a = g["args"]
    leaf = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v) for k, v in g["sd"].items()}
    z, nll = O.listglow_log_prob(g["x"], g["cond"], g["base"], leaf, a["L"], a["K"], a["n_bits"],
                                 noise=g["noise"], learn_prior=True, training=True, **_kw(a))
    close(z.detach(), g["z_logprob"], atol=1e-4)
    close(nll.detach(), g["nll"], rtol=1e-5, atol=1e-3)
    ((nll * g["wts"]).sum() / (0.6931471805599453 * 256 * z.shape[0])).backward()
    for name, ref in g["grads"].items():
        got = leaf[name].grad
        assert got is not None, name
        scale = float(ref.abs().max()) + 1e-12
        assert float((got - ref).abs().max()) <= 2e-4 * scale + 1e-7, name
