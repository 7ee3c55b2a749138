"""Training path of the ConvLSTM (Utils/modules.py:326-414 of the reference): forward that keeps what the backward needs,
and back-propagation through time on librfk's kernels.

Per step t the forward keeps the packed NHWC bf16 input [x_t | h_{t-1}] and the gate pre-activations cc_t (fp32, bias
included); gates are recomputed in the backward.  Backward, for t = T-1 .. 0:

  rfk_convlstm_pointwise_bwd   (dh_t + dh from step t+1, dc from step t+1) -> dcc_t, dc_{t-1}, d bias
  rfk_conv_wgrad               d weight += [x_t | h_{t-1}]^T dcc_t            (tcgen05, accumulates over t)
  rfk_conv_gemm (flipped W)    d[x_t | h_{t-1}] = dcc_t * W^T                 -> dx_t and the dh handed to step t-1

The peephole tensors are constants, as in the reference, so they get no gradient.
"""
import torch

from .. import ops


def _forward(cell, x, ht, ct):
    B, T, C, H, W = x.shape
    hc, dev = cell.hidden_channels, x.device
    conv = cell.conv[0]
    if not cell.init_done:
        cell.initialize_peephole(H, W, dev)
        cell.init_done = True
    cin = C + hc
    wgt, cin_pad = cell._weights_plain()
    bias = None if conv.bias is None else conv.bias.detach().float()
    out = torch.empty(B, T, hc, H, W, device=dev, dtype=torch.float32)
    saved = []
    h_prev, c_prev = ht, ct
    # the packed inputs [x_t | h_{t-1}] of all steps in ONE buffer, time-major: the weight gradient, a reduction over
    # pixels, then runs once over T*B*H*W pixels instead of once per step
    bufs = torch.zeros(T * B, H, W, ops.cin_pad(cin), device=dev, dtype=torch.bfloat16)
    for t in range(T):
        buf = bufs[t * B:(t + 1) * B]
        ops.pack_nhwc(x[:, t], 0, C, buf, 0)
        if h_prev is not None:
            ops.pack_nhwc(h_prev, 0, hc, buf, C)
        cc = torch.empty(B, 4 * hc, H, W, device=dev, dtype=torch.float32)
        if ops.use_small_gemm(B, H, W, cell.taps, cin_pad):
            ops.conv_gemm_small(buf, cin_pad, wgt, 4 * hc, cell.taps, bias, cc)
        else:
            ops.conv_gemm(buf, cin_pad, wgt, 4 * hc, cell.taps, None, bias, "none", cc)
        cp = c_prev if c_prev is not None else torch.zeros(B, hc, H, W, device=dev, dtype=torch.float32)
        h, c = ops.convlstm_pointwise(cc, cp, cell._peep)
        out[:, t].copy_(h)
        saved.append((buf, cc, c_prev))
        h_prev, c_prev = h, c
    return out, c_prev, (saved, bufs)


class _ConvLSTMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cell, x, ht, ct, weight, bias):
        x = ops.f32c(x)
        ht = None if ht is None else ops.f32c(ht)
        ct = None if ct is None else ops.f32c(ct)
        out, c_last, saved = _forward(cell, x, ht, ct)
        ctx.cell, ctx.saved, ctx.shape = cell, saved, x.shape
        ctx.has_state = ht is not None
        ctx.has_bias = bias is not None
        return out, c_last

    @staticmethod
    def backward(ctx, d_out, d_c):
        cell = ctx.cell
        if ctx.saved is None:
            raise RuntimeError("recurrent-flows-msc_b200: ConvLSTM's saved activations were already consumed (no retain_graph)")
        saved, bufs = ctx.saved
        ctx.saved = None
        B, T, C, H, W = ctx.shape
        hc, dev = cell.hidden_channels, d_out.device
        conv = cell.conv[0]
        cin = C + hc
        d_out = ops.f32c(d_out)
        dc = None if d_c is None else ops.f32c(d_c)
        k = 3 if cell.taps == 9 else 1
        dw = torch.zeros(4 * hc, cin, k, k, device=dev, dtype=torch.float32)
        dbias = torch.zeros(4 * hc, device=dev, dtype=torch.float32) if ctx.has_bias else None
        cache = cell.__dict__.setdefault("_wd_cache", [None, None])   # data-gradient weights, rebuilt when the weight changes
        key = (conv.weight.data_ptr(), conv.weight._version, _epoch())
        if cache[0] != key:
            cache[0], cache[1] = key, ops.pack_dgrad_weight(conv.weight)
        need_dx = ctx.needs_input_grad[1]
        if need_dx:
            wd, cp = cache[1]
            n_rows, h_lo = cin, C
        else:
            # the input sequence needs no gradient (e.g. detached features): only the h_{t-1} rows of the data gradient
            hkey = key + ("h",)
            if cache[0] != key or cell.__dict__.get("_wdh_cache", (None,))[0] != hkey:
                rows = torch.arange(C, cin, device=dev)
                cell.__dict__["_wdh_cache"] = (hkey, ops.pack_dgrad_weight(conv.weight, rows))
            wd, cp = cell.__dict__["_wdh_cache"][1]
            n_rows, h_lo = hc, 0
        dx = torch.empty(B, T, C, H, W, device=dev, dtype=torch.float32) if need_dx else None
        das = torch.zeros(T * B, H, W, ops.cin_pad(4 * hc), device=dev, dtype=torch.bfloat16)
        dh_future = None
        for t in reversed(range(T)):
            buf, cc, c_prev = saved[t]
            dh = d_out[:, t]
            if dh_future is not None:
                dh = dh + dh_future
            dcc, dc = ops.convlstm_pointwise_bwd(cc, c_prev, cell._peep, dh, dc, dbias)
            da = das[t * B:(t + 1) * B]
            ops.pack_nhwc(dcc, 0, 4 * hc, da, 0)
            if t == 0 and not need_dx and not ctx.has_state:
                break       # nothing upstream of the first step needs a data gradient
            din = torch.empty(B, n_rows, H, W, device=dev, dtype=torch.float32)
            if ops.use_small_gemm(B, H, W, cell.taps, cp):
                ops.conv_gemm_small(da, cp, wd, n_rows, cell.taps, None, din)
            else:
                ops.conv_gemm(da, cp, wd, n_rows, cell.taps, None, None, "none", din)
            if need_dx:
                dx[:, t].copy_(din[:, :C])
            dh_future = din[:, h_lo:]
        # d weight = sum_t [x_t | h_{t-1}]^T dcc_t: one launch over all T*B*H*W pixels
        ops.conv_wgrad(bufs, cin, das, 4 * hc, cell.taps, out=dw)
        dweight = dw
        d_h0 = dh_future.contiguous() if ctx.has_state else None
        d_c0 = dc if ctx.has_state else None
        return None, dx, d_h0, d_c0, dweight, dbias


def _epoch():
    from ..Flow.glow_modules import _PARAM_EPOCH
    return _PARAM_EPOCH[0]


def convlstm_with_grad(cell, x, ht, ct):
    """(stack of h [B,T,Hc,H,W], h_T, c_T) with gradients for x, the initial state and the gate convolution."""
    conv = cell.conv[0]
    out, c_last = _ConvLSTMFn.apply(cell, x, ht, ct, conv.weight, conv.bias)
    return out, out[:, x.shape[1] - 1], c_last
