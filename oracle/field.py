"""torch-CPU restatement of the D-NeRF field (dnerf/network.py:123-208) and of FFMLP's semantics
(testing/test_ffmlp.py:11-43 torch `MLP`), differentiable with autograd.  TEST INFRASTRUCTURE ONLY.

`half=True` emulates the reference's autocast numerics (operands and layer outputs rounded to fp16, fp32 accumulate)
with straight-through rounding so gradients still flow; `half=False` is the plain fp32 "truth".
"""
import numpy as np
import torch

from . import grid as og

PRIMES = [1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737]


def _q(x, half):
    """Round to fp16 (straight-through gradient)."""
    if not half:
        return x
    return x + (x.detach().half().to(x.dtype) - x.detach())


def freq_encode(x, degree):
    """freqencoder.cu:30-58: [x, sin(2^f x), sin(2^f x + pi/2)] per frequency, channel c -> (col = c/D - 1, d = c % D)."""
    D = x.shape[1]
    outs = [x]
    for f in range(degree):
        outs.append(torch.sin(x * (2.0 ** f)))
        outs.append(torch.sin(x * (2.0 ** f) + np.float32(np.pi / 2)))
    return torch.cat(outs, dim=1)


def sh_encode(d):
    """shencoder.cu:49-70, degree 4."""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    return torch.stack([
        torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2, 0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0), 0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3.0 * y2)], dim=1)


class _TruncExp(torch.autograd.Function):
    """activation.py:5-17."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


def grid_encode(x01, table, offsets, S, H, gridtype=0, align_corners=False, scales=None):
    """Differentiable (w.r.t. x01 and table) restatement of gridencoder.cu:88-197 (linear interpolation).
    Index math is done in int64 with explicit uint32 wrap-around."""
    B, D = x01.shape
    L = len(offsets) - 1
    sc, res = og.level_params(L, S, H, scales)
    outs = []
    M32 = 0xFFFFFFFF
    oob = ((x01 < 0) | (x01 > 1)).any(1)
    for l in range(L):
        hm = int(offsets[l + 1] - offsets[l])
        scale = float(sc[l])
        pos = x01 * scale + (0.0 if align_corners else 0.5)
        pg = torch.floor(pos.detach())
        frac = pos - pg
        pg = pg.long()
        step = int(res[l]) if align_corners else int(res[l]) + 1
        acc = 0
        for idx in range(1 << D):
            w = torch.ones(B, dtype=x01.dtype)
            stride, index, hashed = 1, torch.zeros(B, dtype=torch.long), False
            cs = []
            for d in range(D):
                bit = (idx >> d) & 1
                w = w * (frac[:, d] if bit else (1 - frac[:, d]))
                cs.append((pg[:, d] + bit) & M32)
            for d in range(D):
                if stride > hm:
                    break
                index = (index + cs[d] * stride) & M32
                stride = (stride * step) & M32
            if gridtype == 0 and stride > hm:
                index = torch.zeros(B, dtype=torch.long)
                for d in range(D):
                    index = index ^ ((cs[d] * PRIMES[d]) & M32)
            rows = (index % hm) + int(offsets[l])
            acc = acc + w.unsqueeze(1) * table[rows]
        outs.append(acc)
    out = torch.stack(outs, dim=1).reshape(B, -1)
    return torch.where(oob.unsqueeze(1), torch.zeros_like(out), out)


def mlp(x, weights, half=True):
    """Bias-free MLP with ReLU between layers (nn.Linear weights [out, in])."""
    h = x
    for i, w in enumerate(weights):
        h = _q(h, half) @ _q(w, half).t()
        h = _q(h, half)
        if i != len(weights) - 1:
            h = torch.relu(h)
    return h


def dnerf_forward(xyz, dirs, t, deform_w, sigma_w, color_w, table, offsets, S, H, bound=1.0, density_scale=1.0, half=True,
                  t0_mode=1, scales=None, gridtype=0, deform_values=None):
    """dnerf/network.py:123-169 (t0_mode=1) / :171-208 density (t0_mode=2).  t: python float.
    deform_values (optional): use these numbers for the deformation while keeping the oracle's own gradient path
    (lets a test evaluate the grid at exactly the positions another implementation produced)."""
    tt = torch.full((xyz.shape[0], 1), float(t), dtype=xyz.dtype)
    enc = torch.cat([freq_encode(xyz, 10), freq_encode(tt, 6)], dim=1)
    deform = mlp(enc, deform_w, half)
    if deform_values is not None:
        deform = deform + (deform_values - deform).detach()
    if float(t) == 0.0:
        x = xyz
        if t0_mode == 1:
            deform = torch.zeros_like(xyz)
    else:
        x = xyz + deform
    x01 = (x + bound) / (2 * bound)
    feat = _q(grid_encode(x01, _q(table, half), offsets, S, H, gridtype, False, scales), half)
    h = mlp(feat, sigma_w, half)
    sigma = trunc_exp(h[:, 0]) * density_scale
    geo = h[:, 1:]
    cin = torch.cat([_q(sh_encode(dirs), half), geo], dim=1)
    rgb = _q(torch.sigmoid(mlp(cin, color_w, half)), half)
    return sigma, rgb, deform
