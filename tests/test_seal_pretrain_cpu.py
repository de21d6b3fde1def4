"""Host logic of Seal's local pre-training (seald_nerf_b200/SealDNeRF/pretrain.py) on CPU: the point / direction lattices of
`sample_points` against the reference's construction (torch.arange lattices + scipy's Rotation.from_euler('xyz', ., degrees=True)
applied to (1 - 1e-5, 0, 0), SealDNeRF/utils.py:308-335) and the batch boundaries of `init_pretraining` (:443-447)."""
import numpy as np
import pytest
import torch

from seald_nerf_b200.SealDNeRF.pretrain import SealPretrainer, sample_points


def test_sample_points_matches_reference_construction():
    Rotation = pytest.importorskip("scipy.spatial.transform").Rotation
    bounds = torch.tensor([[[-0.1, 0.0, 0.05], [0.2, 0.15, 0.3]], [[0.0, 0.0, 0.0], [0.05, 0.05, 0.05]]])
    pts, dirs = sample_points(bounds, point_step=0.05, angle_step=45)
    ref_p, ref_d = [], []
    for i in range(2):
        lo, hi = bounds[i]
        X, Y, Z = torch.meshgrid(torch.arange(lo[0], hi[0], step=0.05), torch.arange(lo[1], hi[1], step=0.05), torch.arange(lo[2], hi[2], step=0.05),
                                 indexing="ij")
        ref_p.append(torch.stack([X, Y, Z], dim=-1).reshape(-1, 3))
        r_x, r_y, r_z = torch.meshgrid(torch.arange(0, 360, step=45), torch.arange(0, 360, step=45), torch.arange(0, 360, step=45), indexing="ij")
        eul = torch.stack([r_x, r_y, r_z], dim=-1).reshape(-1, 3)
        ref_d.append(torch.from_numpy(Rotation.from_euler("xyz", eul.numpy(), degrees=True).apply(np.array([1 - 1e-5, 0, 0]))))
    assert torch.equal(pts, torch.concat(ref_p))
    np.testing.assert_allclose(dirs.numpy(), torch.concat(ref_d).numpy(), rtol=0, atol=1e-12)
    assert dirs.shape == (2 * 8 ** 3, 3) and dirs.dtype == torch.float64


def test_single_bound_and_step_lists():
    pts, dirs = sample_points(torch.tensor([[0.0, 0.0, 0.0], [0.35, 0.25, 0.15]]), point_step=0.1, angle_step=90)
    assert pts.shape == (4 * 3 * 2, 3) and dirs.shape == (64, 3)
    assert SealPretrainer._steps(10000, 4096) == [0, 4096, 8192, 10000]
    assert SealPretrainer._steps(8192, 4096) == [0, 4096, 8192]
    assert SealPretrainer._steps(5, 4096) == [0, 5]
    assert SealPretrainer._steps(0, 4096) == [0]
