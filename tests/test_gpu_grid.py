"""GPU parity of the grid encoder against the numpy oracle (oracle/grid.py) and the reference's own extension.

Bar: corner indices, per-level scale and resolution BIT-EXACT; values: fp32 table rtol 1e-4 / atol 1e-6 (of the value
scale), fp16 table rtol 2e-2 / atol 1e-3 * max|value| (fp16 storage + the reference's fp16 accumulation);
table gradients compared after fp32 up-cast with atol = 1e-3 * max|g| (half2 atomics are order dependent).
"""
import numpy as np
import pytest
import torch

from conftest import load_ref

pytestmark = pytest.mark.gpu


def _cfg(D, L, C, log2_T, base, desired, align=False):
    from oracle import grid as og
    offsets, pls = og.make_offsets(D, L, C, 2.0, base, log2_T, desired, align)
    return offsets, float(np.log2(pls)), pls


def _device_scales(dev, offsets, D, L, S, base, gridtype, align):
    """Per-level (scale, resolution) as the device computes them (ex2.approx): within 2 ulp of the libm oracle."""
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    from oracle import grid as og
    tx = torch.zeros(1, D, device=dev); to = torch.from_numpy(offsets).to(dev)
    idx = torch.empty(1, L, 1 << D, dtype=torch.int32, device=dev)
    scales = torch.empty(L, device=dev); ress = torch.empty(L, dtype=torch.int32, device=dev)
    _lib.call("seald_grid_debug_indices", ptr(tx), ptr(to), ptr(idx), ptr(scales), ptr(ress), 1, D, L, S, base, gridtype, int(align),
              _lib.stream())
    sc = scales.cpu().numpy()
    sc_o, _ = og.level_params(L, S, base)
    ulp = np.spacing(np.abs(sc_o) + 1.0)  # scale + 1 = exp2f(.) * H
    assert np.all(np.abs(sc.astype(np.float64) - sc_o) <= 2 * ulp), "device exp2f must stay within 2 ulp of libm"
    assert sc[0] == sc_o[0]  # level 0 is exact
    return sc, ress.cpu().numpy().astype(np.uint32)


def _points(B, D, seed):
    rng = np.random.default_rng(seed)
    x = rng.random((B, D), dtype=np.float32)
    x[0] = 0.0
    x[1] = 1.0
    x[2, 0] = -1e-6  # just outside -> zero row
    x[3, D - 1] = 1.0 + 1e-6
    x[4] = x[5]  # duplicate point (atomic accumulation)
    return x


CASES = [
    # D, L, C, log2_T, base, desired, gridtype, align, interp
    (3, 16, 2, 19, 16, 2048, 0, False, 0),
    (3, 16, 2, 19, 16, 2048, 1, False, 0),
    (4, 16, 2, 19, 16, 2048, 0, False, 0),
    (4, 8, 2, 15, 8, 256, 1, False, 0),
    (2, 4, 2, 19, 16, 2048, 0, False, 0),
    (3, 8, 4, 14, 8, 512, 0, True, 0),
    (3, 8, 1, 14, 8, 512, 0, False, 1),
    (3, 6, 8, 12, 4, 64, 1, True, 1),
]


@pytest.mark.parametrize("case", CASES)
def test_indices_bit_exact(cuda_dev, case):
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    from oracle import grid as og
    D, L, C, log2_T, base, desired, gridtype, align, interp = case
    offsets, S, _ = _cfg(D, L, C, log2_T, base, desired, align)
    B = 3001
    x = _points(B, D, 0)
    x = np.clip(x, 0, 1)  # indices are only defined for in-range points
    d = cuda_dev
    tx = torch.from_numpy(x).to(d); to = torch.from_numpy(offsets).to(d)
    idx = torch.empty(B, L, 1 << D, dtype=torch.int32, device=d)
    scales = torch.empty(L, device=d); ress = torch.empty(L, dtype=torch.int32, device=d)
    _lib.call("seald_grid_debug_indices", ptr(tx), ptr(to), ptr(idx), ptr(scales), ptr(ress), B, D, L, S, base, gridtype, int(align),
              _lib.stream())
    sc_dev, rs_dev = _device_scales(d, offsets, D, L, S, base, gridtype, align)
    assert np.array_equal(scales.cpu().numpy(), sc_dev)
    sc_o, rs_o = og.level_params(L, S, base, sc_dev)
    assert np.array_equal(ress.cpu().numpy().astype(np.uint32), rs_o)
    table = np.zeros((int(offsets[-1]), 1), np.float32)
    _, idx_o = og.grid_encode_forward(x, table, offsets, S, base, gridtype, align, 0, want_indices=True, scales=sc_dev)
    assert np.array_equal(idx.cpu().numpy().astype(np.uint32), idx_o)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("case", CASES)
def test_forward_backward_vs_oracle_and_reference(cuda_dev, case, dtype):
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr, F16, F32
    from oracle import grid as og
    D, L, C, log2_T, base, desired, gridtype, align, interp = case
    if dtype == torch.float16 and C % 2:
        pytest.skip("the reference only uses a half table when C is even (grid.py:43)")
    offsets, S, _ = _cfg(D, L, C, log2_T, base, desired, align)
    B = 4099  # not a multiple of 256/512
    x = _points(B, D, 1)
    rng = np.random.default_rng(2)
    table = (rng.standard_normal((int(offsets[-1]), C)) * 0.1).astype(np.float32)
    d = cuda_dev
    tt = torch.from_numpy(table).to(d).to(dtype)
    table_q = tt.float().cpu().numpy()  # what the device actually holds
    tx = torch.from_numpy(x).to(d); to = torch.from_numpy(offsets).to(d)
    dt = F16 if dtype == torch.float16 else F32
    out = torch.empty(B, L * C, dtype=dtype, device=d)
    dy = torch.empty(B, L, D, C, dtype=dtype, device=d)
    _lib.call("seald_grid_encode_forward", ptr(tx), ptr(tt), ptr(to), ptr(out), ptr(dy), B, D, C, L, S, base, gridtype, int(align),
              interp, dt, None, _lib.stream())
    sc_dev, _ = _device_scales(d, offsets, D, L, S, base, gridtype, align)
    out_o, dy_o = og.grid_encode_forward(x, table_q, offsets, S, base, gridtype, align, interp, want_dy_dx=True, scales=sc_dev)
    if dtype == torch.float32:
        rtol, atol, dy_atol = 1e-4, 1e-6, 1e-3
    else:
        rtol, atol, dy_atol = 2e-2, 1e-3 * np.abs(out_o).max(), 2e-2 * np.abs(dy_o).max()
    np.testing.assert_allclose(out.float().cpu().numpy(), out_o, rtol=rtol, atol=atol)
    np.testing.assert_allclose(dy.float().cpu().numpy(), dy_o, rtol=2e-2 if dtype == torch.float16 else 1e-3, atol=dy_atol)
    assert float(out[2].abs().max()) == 0 and float(out[3].abs().max()) == 0  # out-of-range rows are zero

    # without dy_dx the output is identical
    out2 = torch.empty_like(out)
    _lib.call("seald_grid_encode_forward", ptr(tx), ptr(tt), ptr(to), ptr(out2), None, B, D, C, L, S, base, gridtype, int(align),
              interp, dt, None, _lib.stream())
    assert torch.equal(out, out2)

    # backward: table gradient (dtype of the table, like the reference) + recomputed input gradient
    g = (rng.standard_normal((B, L * C))).astype(np.float32)
    tg = torch.from_numpy(g).to(d).to(dtype)
    g_q = tg.float().cpu().numpy()
    gt_o, gx_o = og.grid_encode_backward(g_q, x, table_q, offsets, S, base, gridtype, align, interp, want_grad_x=True, scales=sc_dev)
    for gdt, gtorch in ((dt, dtype), (F32, torch.float32)):
        grad_table = torch.zeros(int(offsets[-1]), C, dtype=gtorch, device=d)
        grad_x = torch.empty(B, D, device=d)
        _lib.call("seald_grid_encode_backward", ptr(tg), ptr(tx), ptr(tt), ptr(to), ptr(grad_table), None, ptr(grad_x), B, D, C, L, S,
                  base, gridtype, int(align), interp, dt, gdt, None, _lib.stream())
        # an fp16 gradient table is accumulated with half2 atomics (like the reference, gridencoder.cu:325-331): every add
        # rounds to fp16, so cells that collect many points carry an error of a few 1e-3 of the largest entry
        tol = (5e-3 if gtorch == torch.float16 else 1e-5) * np.abs(gt_o).max()
        np.testing.assert_allclose(grad_table.float().cpu().numpy(), gt_o, rtol=3e-2 if gtorch == torch.float16 else 1e-4, atol=tol)
        np.testing.assert_allclose(grad_x.cpu().numpy(), gx_o, rtol=2e-2 if dtype == torch.float16 else 1e-3,
                                   atol=(2e-3 if dtype == torch.float16 else 1e-4) * np.abs(gx_o).max())
    # dy_dx-based input gradient path agrees with the recompute path
    grad_table = torch.zeros(int(offsets[-1]), C, dtype=torch.float32, device=d)
    grad_x2 = torch.empty(B, D, device=d)
    _lib.call("seald_grid_encode_backward", ptr(tg), ptr(tx), ptr(tt), ptr(to), ptr(grad_table), ptr(dy), ptr(grad_x2), B, D, C, L, S,
              base, gridtype, int(align), interp, dt, F32, None, _lib.stream())
    np.testing.assert_allclose(grad_x2.cpu().numpy(), gx_o, rtol=5e-2 if dtype == torch.float16 else 1e-3,
                               atol=(1e-2 if dtype == torch.float16 else 1e-4) * np.abs(gx_o).max())

    ref = load_ref("gridencoder")
    if ref is not None:
        out_r = torch.empty(L, B, C, dtype=dtype, device=d)
        dy_r = torch.empty(B, L * D * C, dtype=dtype, device=d)
        ref.grid_encode_forward(tx, tt, to, out_r, B, D, C, L, S, base, dy_r, gridtype, align, interp)
        out_r = out_r.permute(1, 0, 2).reshape(B, L * C)
        if dtype == torch.float32:
            torch.testing.assert_close(out, out_r, rtol=1e-5, atol=1e-7)
            torch.testing.assert_close(dy.view(B, -1), dy_r, rtol=1e-4, atol=1e-4)
        else:
            torch.testing.assert_close(out.float(), out_r.float(), rtol=2e-2, atol=float(atol))
        gr = tg.view(B, L, C).permute(1, 0, 2).contiguous()
        gt_r = torch.zeros(int(offsets[-1]), C, dtype=dtype, device=d)
        gx_r = torch.zeros(B, D, dtype=dtype, device=d)
        ref.grid_encode_backward(gr, tx, tt, to, gt_r, B, D, C, L, S, base, dy_r, gx_r, gridtype, align, interp)
        mine = torch.zeros(int(offsets[-1]), C, dtype=dtype, device=d)
        _lib.call("seald_grid_encode_backward", ptr(tg), ptr(tx), ptr(tt), ptr(to), ptr(mine), None, None, B, D, C, L, S, base,
                  gridtype, int(align), interp, dt, dt, None, _lib.stream())
        scale = float(gt_r.float().abs().max())
        # the reference accumulates every contribution with an fp16 atomicAdd (order dependent, error grows with the
        # number of points per cell); ours accumulates in fp32 first, so the fp16 comparison gets a wider band
        torch.testing.assert_close(mine.float(), gt_r.float(), rtol=3e-2 if dtype == torch.float16 else 1e-4,
                                   atol=(5e-3 if dtype == torch.float16 else 1e-5) * scale)


def test_grid_encoder_module_autograd(cuda_dev):
    """GridEncoder module: forward under autocast returns fp16 [.., 32]; gradients reach embeddings and inputs."""
    from seald_nerf_b200.gridencoder import GridEncoder
    from oracle import grid as og
    torch.manual_seed(0)
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048).to(cuda_dev)
    enc.embeddings.data.normal_(0, 0.1)
    x = (torch.rand(1000, 3, device=cuda_dev) * 2 - 1).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        y = enc(x, bound=1)
    assert y.dtype == torch.float16 and y.shape == (1000, 32)
    w = torch.randn_like(y, dtype=torch.float32)
    (y.float() * w).sum().backward()
    assert enc.embeddings.grad is not None and enc.embeddings.grad.dtype == torch.float32
    assert x.grad is not None and x.grad.shape == (1000, 3)
    x01 = ((x.detach() + 1) / 2).cpu().numpy()
    tab = enc.embeddings.detach().half().float().cpu().numpy()
    S = float(np.log2(enc.per_level_scale))
    sc_dev, _ = _device_scales(cuda_dev, enc.offsets.cpu().numpy(), 3, 16, S, 16, 0, False)
    out_o = og.grid_encode_forward(x01, tab, enc.offsets.cpu().numpy(), S, 16, scales=sc_dev)
    np.testing.assert_allclose(y.detach().float().cpu().numpy(), out_o, rtol=2e-2, atol=1e-3 * np.abs(out_o).max())
    gt_o, gx_o = og.grid_encode_backward(w.half().float().cpu().numpy(), x01, tab, enc.offsets.cpu().numpy(), S, 16, want_grad_x=True,
                                         scales=sc_dev)
    np.testing.assert_allclose(enc.embeddings.grad.cpu().numpy(), gt_o, rtol=2e-2, atol=2e-3 * np.abs(gt_o).max())
    np.testing.assert_allclose(x.grad.cpu().numpy(), gx_o / 2, rtol=3e-2, atol=3e-3 * np.abs(gx_o).max())  # d x01 / d x = 1/2
    # fp32 path (no autocast) follows the reference's gradcheck-style test (testing/test_hashgrid_grad.py) in spirit
    enc.zero_grad()
    y32 = enc(x.detach(), bound=1)
    assert y32.dtype == torch.float32


def test_split_backward_equals_combined_and_overflow_flag(cuda_dev):
    """seald_grid_encode_backward_table + _input == seald_grid_encode_backward (same kernels, separate launches), and the
    overflow flag of the table scatter is raised exactly when the table gradient is non-finite (GradScaler's found_inf)."""
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr, F16, F32
    D, L, C, log2_T, base, desired = 3, 16, 2, 19, 16, 2048
    offsets, S, _ = _cfg(D, L, C, log2_T, base, desired)
    d = cuda_dev
    B = 5000
    rng = np.random.default_rng(5)
    tx = torch.from_numpy(_points(B, D, 3)).to(d)
    to = torch.from_numpy(offsets).to(d)
    tt = torch.from_numpy((rng.standard_normal((int(offsets[-1]), C)) * 0.1).astype(np.float32)).to(d).half()
    tg = torch.from_numpy(rng.standard_normal((B, L * C)).astype(np.float32)).to(d).half()
    ga, gxa = torch.zeros(int(offsets[-1]), C, device=d), torch.empty(B, D, device=d)
    _lib.call("seald_grid_encode_backward", ptr(tg), ptr(tx), ptr(tt), ptr(to), ptr(ga), None, ptr(gxa), B, D, C, L, S, base, 0, 0, 0, F16,
              F32, None, _lib.stream())
    gb, gxb = torch.zeros_like(ga), torch.empty_like(gxa)
    flag = torch.zeros(1, dtype=torch.int32, device=d)
    _lib.call("seald_grid_encode_backward_table", ptr(tg), ptr(tx), ptr(to), ptr(gb), B, D, C, L, S, base, 0, 0, 0, F16, F32, None,
              ptr(flag), _lib.stream())
    _lib.call("seald_grid_encode_backward_input", ptr(tg), ptr(tx), ptr(tt), ptr(to), None, ptr(gxb), B, D, C, L, S, base, 0, 0, 0, F16,
              None, _lib.stream())
    assert torch.equal(gxa, gxb)
    torch.testing.assert_close(ga, gb, rtol=1e-5, atol=1e-5 * float(ga.abs().max()))  # (atomic order)
    assert int(flag) == 0 and bool(torch.isfinite(gb).all())
    # live-row count: rows >= *b_dev are not consumed, an inf there must not raise the flag ...
    tg2 = tg.clone()
    tg2[4000, 7] = float("inf")
    b_dev = torch.tensor([3000], dtype=torch.int32, device=d)
    gb.zero_()
    _lib.call("seald_grid_encode_backward_table", ptr(tg2), ptr(tx), ptr(to), ptr(gb), B, D, C, L, S, base, 0, 0, 0, F16, F32, ptr(b_dev),
              ptr(flag), _lib.stream())
    assert int(flag) == 0 and bool(torch.isfinite(gb).all())
    # ... and one in a consumed row must, together with a non-finite table gradient
    gb.zero_()
    _lib.call("seald_grid_encode_backward_table", ptr(tg2), ptr(tx), ptr(to), ptr(gb), B, D, C, L, S, base, 0, 0, 0, F16, F32, None,
              ptr(flag), _lib.stream())
    assert int(flag) != 0 and not bool(torch.isfinite(gb).all())
    # an out-of-range point contributes nothing (gridencoder.cu:276-281): its inf is not consumed either
    tg3 = tg.clone()
    tg3[2, 0] = float("nan")  # _points() puts row 2 just outside [0,1]
    flag.zero_(); gb.zero_()
    _lib.call("seald_grid_encode_backward_table", ptr(tg3), ptr(tx), ptr(to), ptr(gb), B, D, C, L, S, base, 0, 0, 0, F16, F32, None,
              ptr(flag), _lib.stream())
    assert int(flag) == 0 and bool(torch.isfinite(gb).all())
