"""ENet encoder as a flat layer program with BatchNorm folded away (host side; reference:
lib/nets/enet.py:130-590, create_enet_for_3d :697-715).

`compile_enet(params)` turns the reference state_dict (tensors in its own order, num_batches_tracked dropped) into a
list of ops a conv kernel with a fused epilogue can execute:

    ("conv", w[Cout,Cin,kh,kw], bias[Cout], stride, (ph,pw), dilation, slope[Cout] | None, src, dst)
        dst = prelu(conv(src) + bias, slope)            (slope None: no activation)
    ("pool", src, dst)                                   2x2/2 max-pool
    ("affine_prelu", scale[C], shift[C], slope[C], src, dst)   per-channel affine + PReLU (initial block's BatchNorm acts
                                                         on concat(conv, pool): the pooled channels have no conv to fold into)
    ("add_prelu", a, b, slope[C], dst, pad_channels)     dst = prelu(a + zero_pad_channels(b), slope)

Every eval-mode BatchNorm that follows a convolution is folded into that convolution's weights and bias (w' = w * g /
sqrt(var + eps), b' = (b - mean) * g / sqrt(var + eps) + beta), and the inference-time Dropout2d factor (1 - p) of the
reference's Torch7-style Dropout (enet.py:89-95) into the bottleneck's last 1x1 conv.  `run_program` executes the program
with torch CPU ops; tests/test_enet_program.py checks it against the features of the unmodified reference ENet."""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-3
_STAGE23 = [("reg", 1), ("reg", 2), ("asym", 5), ("reg", 4), ("reg", 1), ("reg", 8), ("asym", 5), ("reg", 16)]
PROGRAM = ([("down", 16, 64, 0.01)] + [("reg", 16, 64, 0.01, 1)] * 4 + [("down", 32, 128, 0.1)] +
           [(kind, 32, 128, 0.1, arg) for kind, arg in _STAGE23 * 2])


def _fold(w, b, bn, extra=1.0):
    g, beta, mean, var = bn
    s = g / torch.sqrt(var + BN_EPS) * extra
    b0 = torch.zeros_like(mean) if b is None else b
    return w * s.view(-1, 1, 1, 1), (b0 - mean) * s + beta * extra


def compile_enet(params):
    it = iter(params)
    take = lambda n: [next(it) for _ in range(n)]
    ops = []
    # initial block: conv 3->13 (k3 s2 p1) || max-pool, concat, BatchNorm, PReLU
    w0, b0 = take(2)
    bn = take(4)
    slope = next(it)
    s = bn[0] / torch.sqrt(bn[3] + BN_EPS)
    ops.append(("conv", w0 * s[:13].view(-1, 1, 1, 1), (b0 - bn[2][:13]) * s[:13] + bn[1][:13], 2, (1, 1), 1, slope[:13], "in", "c0"))
    ops.append(("pool", "in", "p0"))
    ops.append(("affine_prelu", s[13:], bn[1][13:] - bn[2][13:] * s[13:], slope[13:], "p0", "p0a"))
    ops.append(("cat", "c0", "p0a", "x"))
    for op in PROGRAM:
        kind, mid, cout, p = op[0], op[1], op[2], op[3]
        w1 = next(it)
        w, b = _fold(w1, None, take(4))
        ops.append(("conv", w, b, 2 if kind == "down" else 1, (0, 0), 1, next(it), "x", "y"))
        if kind == "asym":
            wa = next(it)                 # 1x5, no bias, no BatchNorm of its own
            wb, bb = take(2)              # 5x1 + bias, then BatchNorm + PReLU
            ops.append(("conv", wa, torch.zeros(wa.shape[0]), 1, (0, 2), 1, None, "y", "y"))
            w, b = _fold(wb, bb, take(4))
            ops.append(("conv", w, b, 1, (2, 0), 1, next(it), "y", "y"))
        else:
            d = op[4] if kind == "reg" else 1
            w2, b2 = take(2)
            w, b = _fold(w2, b2, take(4))
            ops.append(("conv", w, b, 1, (d, d), d, next(it), "y", "y"))
        w3 = next(it)
        w, b = _fold(w3, None, take(4), extra=1.0 - p)  # BatchNorm and the inference Dropout2d factor
        ops.append(("conv", w, b, 1, (0, 0), 1, None, "y", "y"))
        if kind == "down":
            ops.append(("pool", "x", "s"))
            ops.append(("add_prelu", "y", "s", next(it), "x", cout))
        else:
            ops.append(("add_prelu", "y", "x", next(it), "x", cout))
    if next(it, None) is not None:
        raise ValueError("compile_enet: unused parameters (not the encoder's state_dict?)")
    return ops


def run_program(ops, images):
    """Execute a compiled program with torch CPU ops (checker for compile_enet; the CUDA executor replaces this)."""
    t = {"in": images}
    for op in ops:
        if op[0] == "conv":
            _, w, b, stride, pad, dil, slope, src, dst = op
            y = F.conv2d(t[src], w, b, stride=stride, padding=pad, dilation=dil)
            t[dst] = y if slope is None else F.prelu(y, slope)
        elif op[0] == "pool":
            t[op[2]] = F.max_pool2d(t[op[1]], 2, 2)
        elif op[0] == "affine_prelu":
            _, sc, sh, slope, src, dst = op
            t[dst] = F.prelu(t[src] * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1), slope)
        elif op[0] == "cat":
            t[op[3]] = torch.cat((t[op[1]], t[op[2]]), 1)
        elif op[0] == "add_prelu":
            _, a, bname, slope, dst, cout = op
            bsrc = t[bname]
            if bsrc.shape[1] < cout:
                bsrc = torch.cat((bsrc, bsrc.new_zeros(bsrc.shape[0], cout - bsrc.shape[1], *bsrc.shape[2:])), 1)
            t[dst] = F.prelu(t[a] + bsrc, slope)
    return t["x"]
