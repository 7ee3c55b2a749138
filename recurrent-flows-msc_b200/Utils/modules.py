"""ConvLSTM of the reference's ``Utils/modules.py`` on librfk's sm_100a kernels.

Same class names, constructor arguments, call signatures and ``state_dict`` keys
(``LSTMlayer.conv.0.{weight,bias}``) as cdglissov/recurrent-flows-msc.  One cell step is ONE kernel:
the gate convolution over cat[x, h] runs as a tcgen05 implicit GEMM and the i/f/o/g update is its
epilogue (rfk_conv_gemm_lstm); the cell state never leaves fp32.
"""
import torch
import torch.nn as nn

from .. import ops


def _param_epoch():
    from ..Flow.glow_modules import _PARAM_EPOCH   # bumped by invalidate_caches() (raw-pointer parameter updates)
    return _PARAM_EPOCH[0]


def _no_split_training():
    if ops.SPLIT:
        raise NotImplementedError("recurrent-flows-msc_b200: RFK_CONV_PRECISION=bf16x3 covers the forward direction; train in bf16")


class ActFun(nn.Module):
    """Utils/modules.py:8-19.  Inside the fused networks the activation is a conv epilogue flag;
    called on its own it is a plain elementwise op."""

    def __init__(self, non_lin, in_place=False):
        super().__init__()
        if non_lin == 'relu':
            self.net = nn.ReLU(inplace=in_place)
        elif non_lin == 'leakyrelu':
            self.net = nn.LeakyReLU(negative_slope=0.20, inplace=in_place)
        else:
            assert False, 'Please specify a activation type from the set {relu,leakyrelu}'
        self.kind = non_lin

    def forward(self, x):
        return self.net(x)


def _hidden_tiling(hidden, m_tiles=1 << 30):
    """Hidden channels per N tile (ht) and its padding to a multiple of 8 columns per gate (4*ht_pad <= 256).

    Large pixel counts take the widest tile (tcgen05.mma costs the same for any N <= 256, so wide tiles do the
    most work per activation byte).  When the pixel tiles alone cannot fill the GPU (RFN: 120 pixels = 1 tile)
    the hidden channels are cut into many narrow tiles so that more SMs stream the weights in parallel."""
    cands = [d for d in range(1, min(hidden, 64) + 1) if hidden % d == 0]
    exact = [d for d in cands if d % 8 == 0] or cands
    for ht in sorted(exact, reverse=True):
        if m_tiles * (hidden // ht) >= 128:
            return ht, ops.pad_to(ht, 8)
    ht = min(d for d in exact if d >= 8) if any(d >= 8 for d in exact) else max(exact)
    return ht, ops.pad_to(ht, 8)


class ConvLSTMLayer(nn.Module):
    """Utils/modules.py:326-393.  Gate order i, f, o, g; the output gate peeps at c_next."""

    def __init__(self, in_channels, hidden_channels, kernel_size, bias, dropout=0, peephole=True, norm=False):
        super().__init__()
        if norm or dropout != 0:
            raise NotImplementedError("ConvLSTMLayer(norm=True / dropout>0) is unused by every caller in the "
                                      "reference and not part of the B200 hot path")
        self.in_channels = in_channels
        self.hidden_channels = hidden_channels
        self.kernel_size = kernel_size
        self.peephole = peephole
        self.padding = ((kernel_size[0] - 1) // 2, (kernel_size[1] - 1) // 2)
        self.bias = bias
        self.taps = kernel_size[0] * kernel_size[1]
        assert self.taps in (1, 9), "kernel must be 1x1 or 3x3"
        self.conv = nn.Sequential(nn.Conv2d(in_channels + hidden_channels, 4 * hidden_channels, kernel_size,
                                            stride=1, padding=self.padding, bias=bias))
        self.init_done = False
        self.apply(self.initialize_weights)
        self._packed = None
        self._peep = None

    def initialize_weights(self, layer):
        """Utils/modules.py:379-383: xavier-normal weights, bias ~ U(0,1)."""
        if type(layer) == nn.Conv2d:
            nn.init.xavier_normal_(layer.weight)
            if layer.bias is not None:
                nn.init.uniform_(layer.bias)

    def initialize_peephole(self, height, width, device):
        """Utils/modules.py:385-393.  In the reference these end up as constant zero tensors that are
        never trained or saved; they are honoured by the kernel if someone fills them."""
        if self.peephole:
            self._peep = torch.zeros(3, self.hidden_channels, height, width, device=device)
            self.Wci, self.Wcf, self.Wco = (self._peep[i:i + 1] for i in range(3))
        else:
            self._peep = None
            self.Wci = self.Wcf = self.Wco = 0

    def _weights(self, m_tiles=1 << 30):
        conv = self.conv[0]
        hc = self.hidden_channels
        ht, ht_pad = _hidden_tiling(hc, m_tiles)
        key = (conv.weight.data_ptr(), conv.weight._version,
               None if conv.bias is None else (conv.bias.data_ptr(), conv.bias._version), ht, _param_epoch())
        if self._packed is None or self._packed[0] != key:
            dev = conv.weight.device
            # tile-interleaved rows: (tile t, gate g, j) <- reference row g*hc + t*ht + j
            t = torch.arange(hc // ht, device=dev)[:, None, None]
            g = torch.arange(4, device=dev)[None, :, None]
            j = torch.arange(ht_pad, device=dev)[None, None, :]
            rows = (g * hc + t * ht + j).expand(hc // ht, 4, ht_pad).clone()
            rows[:, :, ht:] = -1
            rows = rows.reshape(-1)
            with torch.no_grad():
                wgt, cin_pad = ops.pack_conv_weight(conv.weight, None, rows, n_pad=rows.numel())
                b = None
                if conv.bias is not None:
                    b = torch.zeros(rows.numel(), device=dev)
                    b[rows >= 0] = conv.bias.detach().float()[rows[rows >= 0]]
            self._packed = (key, wgt, cin_pad, b, ht, ht_pad)
        return self._packed[1:]

    def _split_k(self, m_tiles):
        """K slices for launches whose pixel tiles cannot fill the GPU: one slice per filter tap (3x3) when a single
        pixel tile would otherwise walk the whole K = taps*(Cin+Hc) loop on a handful of SMs."""
        if ops.SPLIT:
            return 1      # the split-K path hands h to the next step as one bf16 word; the fused kernel writes hi and lo
        return self.taps if (m_tiles <= 2 and self.taps > 1 and self.in_channels + self.hidden_channels >= 256) else 1

    def _weights_plain(self):
        """Natural row order (i,f,o,g blocks of Hc rows) for the split-K path; cached."""
        conv = self.conv[0]
        key = (conv.weight.data_ptr(), conv.weight._version, _param_epoch())
        hit = self.__dict__.get("_packed_plain")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                wgt, cin_pad = ops.pack_conv_weight(conv.weight)
            hit = (key, wgt, cin_pad)
            self.__dict__["_packed_plain"] = hit
        return hit[1], hit[2]

    def step(self, in_buf, c_cur, h_out, next_buf):
        """One cell step on a packed NHWC bf16 input; returns c_next.  Fused GEMM + cell update in one kernel, or,
        for tiny maps, a split-K GEMM into an fp32 workspace followed by the pointwise cell kernel."""
        B, H, W, _ = in_buf.shape
        k_split = self._split_k((B * H * W + 127) // 128)
        if k_split > 1:
            wgt, cin_pad = self._weights_plain()
            hc = self.hidden_channels
            ws = ops.workspace(("lstm_ws", hc), (B * H * W, wgt.shape[0]), in_buf.device, torch.float32)  # zero-initialised
            c_next = torch.empty(h_out.shape, device=h_out.device, dtype=torch.float32)
            ops.conv_gemm_splitk(in_buf, cin_pad, wgt, 4 * hc, self.taps, k_split, ws)
            bias = self.conv[0].bias
            ops.convlstm_pointwise_ws(ws, None if bias is None else bias.detach(), c_cur, self._peep, h_out, c_next,
                                      next_buf, self.in_channels, zero_cc=True)   # leaves ws zeroed for the next step
            return c_next
        wgt, cin_pad, b, ht, ht_pad = self._weights((B * H * W + 127) // 128)
        c_next = torch.empty(h_out.shape, device=h_out.device, dtype=torch.float32)
        ops.conv_gemm_lstm(in_buf, cin_pad, wgt, self.hidden_channels, ht, ht_pad, self.taps, b, c_cur, self._peep,
                           c_next, h_out, next_buf, self.in_channels)
        return c_next

    def _staging(self, B, H, W, device):
        cin = self.in_channels + self.hidden_channels
        shape = (B, H, W, ops.buf_ld(cin))
        return [ops.workspace(("lstm_in", cin, i), shape, device) for i in range(2)]

    def forward(self, input_tensor, cur_state):
        if torch.is_grad_enabled():
            _no_split_training()
            from .training import convlstm_with_grad   # training: saves pre-activations, hand-written BPTT
            _, h, c = convlstm_with_grad(self, input_tensor.unsqueeze(1), cur_state[0], cur_state[1])
            return h, c
        x = ops.f32c(input_tensor)
        b, c, h, w = x.shape
        if not self.init_done:
            self.initialize_peephole(h, w, x.device)
            self.init_done = True
        buf = self._staging(b, h, w, x.device)[0]
        ops.pack_nhwc(x, 0, c, buf, 0)
        if cur_state[0] is None:
            buf[..., c:c + self.hidden_channels].zero_()
            if ops.SPLIT:
                lo = buf.shape[-1] // 2
                buf[..., lo + c:lo + c + self.hidden_channels].zero_()
            c_cur = None
        else:
            h_cur, c_cur = ops.f32c(cur_state[0]), ops.f32c(cur_state[1])
            ops.pack_nhwc(h_cur, 0, self.hidden_channels, buf, c)
        h_next = torch.empty(b, self.hidden_channels, h, w, device=x.device, dtype=torch.float32)
        c_next = self.step(buf, c_cur, h_next, None)
        return h_next, c_next


class ConvLSTM(nn.Module):
    """Utils/modules.py:396-414: T sequential cell steps; h is handed to the next step as bf16 NHWC by
    the kernel epilogue (ping-pong staging buffers), so each step is pack(x_t) + one fused kernel."""

    def __init__(self, in_channels, hidden_channels, kernel_size, bias=True, dropout=0, peephole=True, norm=False):
        super().__init__()
        self.hidden_channels = hidden_channels
        self.LSTMlayer = ConvLSTMLayer(in_channels=in_channels, hidden_channels=hidden_channels,
                                       kernel_size=kernel_size, bias=bias, dropout=dropout, peephole=peephole,
                                       norm=norm)

    def forward(self, x, ht=None, ct=None):
        if torch.is_grad_enabled():
            _no_split_training()
            from .training import convlstm_with_grad
            return convlstm_with_grad(self.LSTMlayer, x, ht, ct)
        cell = self.LSTMlayer
        x = ops.f32c(x)
        b, seq_len, channel, h, w = x.size()
        hc = self.hidden_channels
        if not cell.init_done:
            cell.initialize_peephole(h, w, x.device)
            cell.init_done = True
        bufs = cell._staging(b, h, w, x.device)
        out = torch.empty(b, seq_len, hc, h, w, device=x.device, dtype=torch.float32)
        if ht is None:
            bufs[0][..., channel:channel + hc].zero_()
            if ops.SPLIT:
                lo = bufs[0].shape[-1] // 2
                bufs[0][..., lo + channel:lo + channel + hc].zero_()
            c_cur = None
        else:
            ops.pack_nhwc(ops.f32c(ht), 0, hc, bufs[0], channel)
            c_cur = ops.f32c(ct)
        for t in range(seq_len):
            cur, nxt = bufs[t % 2], bufs[(t + 1) % 2]
            ops.pack_nhwc(x[:, t], 0, channel, cur, 0)  # batch-strided slice, no copy
            c_cur = cell.step(cur, c_cur, out[:, t], nxt if t + 1 < seq_len else None)
        return out, out[:, seq_len - 1], c_cur
