"""CPU: oracle/occupancy.py against the reference's own expressions (dnerf/renderer.py:477-497, 541-543) evaluated with torch on CPU, and
against the committed golden vectors of the reference's morton3D / packbits kernels (tests/golden/raymarch.npz)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import occupancy as oo


def test_cell_points_equal_reference_expressions():
    H = 128
    for bound in (1, 2):
        torch.manual_seed(bound)
        X = torch.arange(H, dtype=torch.int32)
        xx, yy, zz = torch.meshgrid(X[:40], X[50:70], X[100:128], indexing="ij")  # custom_meshgrid
        coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
        xyzs = 2 * coords.float() / (H - 1) - 1
        half_grid_size = bound / H
        cas_xyzs = xyzs * (bound - half_grid_size)
        u = torch.rand_like(cas_xyzs)
        cas_xyzs += (u * 2 - 1) * half_grid_size
        mine = oo.cell_points(coords.numpy(), u.numpy(), H, bound)
        # torch CPU divides by the scalar exactly, CUDA multiplies by the fp32 reciprocal (what the oracle and the kernel restate): 1 ulp
        np.testing.assert_allclose(mine, cas_xyzs.numpy(), rtol=0, atol=2.5e-7 * bound)


def test_ema_max_equals_reference_masked_update():
    rng = np.random.default_rng(0)
    grid = rng.random(4096).astype(np.float32) * 20
    grid[::7] = -1.0                                   # untrained cells
    tmp = -np.ones(4096, np.float32)
    tmp[rng.integers(0, 4096, 1500)] = rng.random(1500).astype(np.float32) * 30
    g, t = torch.from_numpy(grid.copy()), torch.from_numpy(tmp)
    valid_mask = (g >= 0) & (t >= 0)
    g[valid_mask] = torch.maximum(g[valid_mask] * 0.95, t[valid_mask])  # dnerf/renderer.py:541-543
    assert np.array_equal(oo.ema_max(grid, tmp, 0.95), g.numpy())
    assert np.array_equal(oo.ema_max(grid, tmp, 0.95)[::7], grid[::7])


def test_morton_and_packbits_match_reference_golden():
    """oracle/occupancy.py's morton3D / packbits against the outputs of the reference's own kernels (tests/golden/raymarch.npz, inputs from
    tests/golden/cases.py)."""
    import sys
    sys.path.insert(0, GOLDEN)
    import cases
    g = np.load(os.path.join(GOLDEN, "raymarch.npz"))
    c = cases.raymarch_case()
    assert np.array_equal(oo.morton3D(c["coords"]), g["morton"].astype(np.int64))
    assert np.array_equal(oo.packbits(np.asarray(c["density"]).reshape(-1), c["thresh"]), g["packbits"].reshape(-1))
    grid = np.random.default_rng(1).random(1024).astype(np.float32)
    assert np.array_equal(np.unpackbits(oo.packbits(grid, 0.5), bitorder="little"), (grid > 0.5).astype(np.uint8))
