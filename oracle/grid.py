"""numpy restatement of the reference's grid encoder (gridencoder/src/gridencoder.cu, gridencoder/grid.py).

TEST INFRASTRUCTURE ONLY.  Index math is done in uint32 with wrap-around exactly as on the device
(gridencoder.cu:50-84); positions are fp32 with the device's FMA (x*scale+0.5 is one fused op in the SASS);
interpolation weights are fp32 products in the reference's order; accumulation is float64 (the "truth" the
fp16/fp32 device results are compared against with a stated tolerance).
"""
import numpy as np

PRIMES = np.array([1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737], dtype=np.uint32)


def make_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2.0, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, align_corners=False):
    """grid.py:100-129 -> (offsets int32 [L+1], per_level_scale)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), per_level_scale


def level_params(L, S, H, scales=None):
    """gridencoder.cu:138-139 in fp32: scale = fma(exp2f(level*S), H, -1); resolution = ceil(scale)+1.

    On the GPU exp2f is the hardware ex2.approx (<= 2 ulp from libm), so for non-integer level*S the device value
    of `scale` can differ from this libm-based one in the last bit.  Callers that need bit-exact indices pass the
    per-level scales the device computed (seald_grid_debug_indices / tests/golden/grid_scales.npz) as `scales`.
    """
    S = np.float32(S)
    res = np.empty(L, np.uint32)
    if scales is None:
        scales = np.empty(L, np.float32)
        for l in range(L):
            e = np.exp2(np.float32(np.float32(l) * S)).astype(np.float32)
            scales[l] = np.float32(np.float64(e) * np.float64(H) - 1.0)
    scales = np.asarray(scales, np.float32)
    for l in range(L):
        res[l] = np.uint32(np.ceil(scales[l])) + np.uint32(1)
    return scales, res


def grid_rows(pos_grid, gridtype, align_corners, hashmap_size, resolution):
    """gridencoder.cu:66-84 without the `* C + ch`.  pos_grid: uint32 [B, D]."""
    B, D = pos_grid.shape
    stride = np.uint32(1)
    index = np.zeros(B, np.uint32)
    hm = np.uint32(hashmap_size)
    step = np.uint32(resolution if align_corners else resolution + 1)
    with np.errstate(over="ignore"):
        for d in range(D):
            if stride > hm:
                break
            index = index + pos_grid[:, d] * stride
            stride = np.uint32((int(stride) * int(step)) & 0xFFFFFFFF)
        if gridtype == 0 and stride > hm:
            index = np.zeros(B, np.uint32)
            for d in range(D):
                index = index ^ (pos_grid[:, d] * PRIMES[d])
    return index % hm


def _locate(x01, scale, align_corners, interp):
    x64 = x01.astype(np.float64)
    pos = (x64 * np.float64(scale) + (0.0 if align_corners else 0.5)).astype(np.float32)  # one FMA on the device
    pg = np.floor(pos)
    frac = (pos - pg).astype(np.float32)
    pos_grid = pg.astype(np.int64).astype(np.uint32)
    if interp == 1:
        deriv = (np.float32(6) * frac * (np.float32(1) - frac)).astype(np.float32)
        frac = (frac * frac * (np.float32(3) - np.float32(2) * frac)).astype(np.float32)
    else:
        deriv = np.ones_like(frac)
    return frac, deriv, pos_grid


def _corner_weight(frac, idx, skip=None, start=1.0):
    w = np.full(frac.shape[0], start, np.float32)
    for d in range(frac.shape[1]):
        if d == skip:
            continue
        w = (w * ((frac[:, d]) if (idx >> d) & 1 else (np.float32(1) - frac[:, d]))).astype(np.float32)
    return w


def grid_encode_forward(x01, table, offsets, S, H, gridtype=0, align_corners=False, interp=0, want_dy_dx=False,
                        want_indices=False, scales=None):
    """x01 [B,D] fp32 in [0,1]; table [rows,C] (any float dtype) -> out [B, L*C] float64 (+ dy_dx [B,L,D,C], indices)."""
    x01 = np.ascontiguousarray(x01, np.float32)
    tab = np.asarray(table).astype(np.float64)
    B, D = x01.shape
    L = offsets.shape[0] - 1
    C = tab.shape[1]
    scales, ress = level_params(L, S, H, scales)
    oob = ((x01 < 0) | (x01 > 1)).any(1)
    out = np.zeros((B, L, C))
    dy_dx = np.zeros((B, L, D, C)) if want_dy_dx else None
    indices = np.zeros((B, L, 1 << D), np.uint32) if want_indices else None
    for l in range(L):
        hm = int(offsets[l + 1] - offsets[l])
        frac, deriv, pos_grid = _locate(x01, scales[l], align_corners, interp)
        tl = tab[offsets[l]:offsets[l + 1]]
        vals = []
        for idx in range(1 << D):
            pg = pos_grid.copy()
            for d in range(D):
                if (idx >> d) & 1:
                    pg[:, d] += np.uint32(1)
            rows = grid_rows(pg, gridtype, align_corners, hm, int(ress[l]))
            if want_indices:
                indices[:, l, idx] = rows
            v = tl[rows]
            vals.append(v)
            out[:, l, :] += _corner_weight(frac, idx).astype(np.float64)[:, None] * v
        if want_dy_dx:
            for gd in range(D):
                for idx in range(1 << D):
                    if (idx >> gd) & 1:
                        continue
                    w = _corner_weight(frac, idx, skip=gd, start=scales[l]).astype(np.float64)
                    dy_dx[:, l, gd, :] += (w * deriv[:, gd].astype(np.float64))[:, None] * (vals[idx | (1 << gd)] - vals[idx])
    out[oob] = 0
    if want_dy_dx:
        dy_dx[oob] = 0
    res = [out.reshape(B, L * C)]
    if want_dy_dx:
        res.append(dy_dx)
    if want_indices:
        res.append(indices)
    return res[0] if len(res) == 1 else tuple(res)


def grid_encode_backward(grad_out, x01, table, offsets, S, H, gridtype=0, align_corners=False, interp=0, want_grad_x=False,
                         scales=None):
    """grad_out [B, L*C] -> grad_table [rows, C] float64 (gridencoder.cu:249-340), grad_x [B,D] (gridencoder.cu:344-369)."""
    x01 = np.ascontiguousarray(x01, np.float32)
    B, D = x01.shape
    L = offsets.shape[0] - 1
    C = np.asarray(table).shape[1]
    g = np.asarray(grad_out).astype(np.float64).reshape(B, L, C)
    scales, ress = level_params(L, S, H, scales)
    inb = ~((x01 < 0) | (x01 > 1)).any(1)
    grad_table = np.zeros((int(offsets[-1]), C))
    for l in range(L):
        hm = int(offsets[l + 1] - offsets[l])
        frac, deriv, pos_grid = _locate(x01, scales[l], align_corners, interp)
        for idx in range(1 << D):
            pg = pos_grid.copy()
            for d in range(D):
                if (idx >> d) & 1:
                    pg[:, d] += np.uint32(1)
            rows = grid_rows(pg, gridtype, align_corners, hm, int(ress[l]))
            w = _corner_weight(frac, idx).astype(np.float64)
            np.add.at(grad_table, (rows[inb].astype(np.int64) + int(offsets[l])), (w[:, None] * g[:, l, :])[inb])
    if not want_grad_x:
        return grad_table
    _, dy_dx = grid_encode_forward(x01, table, offsets, S, H, gridtype, align_corners, interp, want_dy_dx=True, scales=scales)
    grad_x = np.einsum("blc,bldc->bd", g, dy_dx)
    return grad_table, grad_x
