"""CPU: oracle/rays.py pinned against tests/golden/rays.npz (outputs of the reference's own get_rays source, see
tests/golden/make_rays_golden.py).  Tolerance: rays_o exact; rays_d rtol 1e-6 / atol 1e-7 (torch evaluates the norm and the
3x3 matmul with its own reduction order)."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import rays as orr


def _cases():
    g = np.load(os.path.join(GOLDEN, "rays.npz"))
    for k in range(2):
        yield k, {n[len("case%d_" % k):]: g[n] for n in g.files if n.startswith("case%d_" % k)}


def test_get_rays_matches_reference_output():
    for k, c in _cases():
        H, W, N = [int(v) for v in c["HWN"]]
        assert c["inds"].shape == (N,) and c["inds"].min() >= 0 and c["inds"].max() < H * W
        ro, rd = orr.get_rays(c["pose"], c["intr"], H, W, c["inds"])
        assert np.array_equal(ro, c["rays_o"])
        np.testing.assert_allclose(rd, c["rays_d"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(np.linalg.norm(rd, axis=-1), 1.0, atol=1e-6)


def test_gather_gt_matches_reference_gather_and_blend():
    for k, c in _cases():
        if c["image"].shape[0] == 0:
            continue
        px = c["image"][c["inds"]]
        assert np.array_equal(px, c["gt_rgba"])  # torch.gather of the flattened image == plain row indexing
        gt = orr.gather_gt(c["image"], c["inds"])
        ref = px[:, :3] * px[:, 3:] + 1.0 * (1 - px[:, 3:])  # dnerf/utils.py:61-66 with bg_color = 1
        np.testing.assert_allclose(gt, ref, rtol=1e-6, atol=1e-7)
        assert np.array_equal(orr.gather_gt(c["image"][:, :3], c["inds"]), px[:, :3])
