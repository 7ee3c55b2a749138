"""Functional fp32 CPU restatement of the reference's ConvLSTM.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import torch
import torch.nn.functional as F


def convlstm_cell(x, h, c, weight, bias, peephole=None):
    """Utils/modules.py:355-377.

    Gate conv over cat[x, h] -> 4*Hc channels split in the order i, f, o, g;
    i = sig(cc_i + Wci*c), f = sig(cc_f + Wcf*c), g = tanh(cc_g),
    c' = f*c + i*g, o = sig(cc_o + Wco*c'), h' = o*tanh(c').
    ``h``/``c`` None means a zero initial state (Utils/modules.py:357-359).
    ``peephole`` is (Wci, Wcf, Wco), each [1,Hc,H,W], or None for zeros
    (Utils/modules.py:385-393: the reference's peepholes start at zero and are never trained).
    """
    hc = weight.shape[0] // 4
    b, _, hh, ww = x.shape
    if h is None:
        h = torch.zeros(b, hc, hh, ww)
        c = torch.zeros(b, hc, hh, ww)
    pad = ((weight.shape[2] - 1) // 2, (weight.shape[3] - 1) // 2)
    cc = F.conv2d(torch.cat([x, h], 1), weight, bias, 1, pad)
    cc_i, cc_f, cc_o, cc_g = torch.split(cc, hc, dim=1)
    if peephole is None:
        wci = wcf = wco = 0.0
    else:
        wci, wcf, wco = peephole
    i = torch.sigmoid(cc_i + wci * c)
    f = torch.sigmoid(cc_f + wcf * c)
    g = torch.tanh(cc_g)
    c_next = f * c + i * g
    o = torch.sigmoid(cc_o + wco * c_next)
    h_next = o * torch.tanh(c_next)
    return h_next, c_next


def convlstm(x, weight, bias, h=None, c=None, peephole=None):
    """Utils/modules.py:406-414.  x [B,T,C,H,W] -> (stack of h [B,T,Hc,H,W], h_T, c_T)."""
    outs = []
    for t in range(x.shape[1]):
        h, c = convlstm_cell(x[:, t], h, c, weight, bias, peephole)
        outs.append(h)
    return torch.stack(outs, 1), h, c
