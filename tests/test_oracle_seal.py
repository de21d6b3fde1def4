"""Pins oracle/seal.py against the REFERENCE's own Seal runtime (SealNeRF/seal_utils.py run on CPU by
tests/golden/make_seal_golden.py -> tests/golden/seal.npz).

Bar: map masks bit-exact (the fixtures keep points off triangle edges by construction: random floats); mapped points /
dirs rtol 1e-5 atol 1e-6 (fp32 matmul order; brush 'linear' atol 2e-5: torch.cdist's expansion); colours atol 2e-6.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import seal_cases  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "seal.npz"))
CASES = seal_cases.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_map_to_origin_matches_reference(name):
    from oracle import seal as S
    mp, pts, dirs, cols = CASES[name]
    p, d, m = S.map_to_origin(mp, pts, dirs)
    assert np.array_equal(m, GOLD[name + "_mask"]), "map mask must be bit-exact"
    # brush 'linear': the reference's torch.cdist uses the |a|^2 + |b|^2 - 2ab expansion in fp32 (abs error ~5e-6 on the distance)
    np.testing.assert_allclose(p, GOLD[name + "_points"], rtol=1e-5, atol=2e-5 if mp["type"] == "brush" else 1e-6)
    np.testing.assert_allclose(d, GOLD[name + "_dirs"], rtol=1e-5, atol=1e-6)
    if name == "bbox_none":
        assert not m.any() and np.array_equal(p, pts)
    else:
        assert m.sum() > 30
        if mp["type"] != "anchor":  # (the anchor mapper's returned mask is its cone filter over ALL points, seal_utils.py:546-551)
            assert not m[-7:].any() and not m[-9]  # zero rows / zero coordinates never map (`points.all(1)`)
    if name + "_colors" in GOLD:
        c = S.map_color(mp, p[m], d[m], cols[m])
        np.testing.assert_allclose(c, GOLD[name + "_colors"], rtol=0, atol=2e-6)


def test_hsv_round_trip():
    from oracle import seal as S
    rng = np.random.default_rng(0)
    rgb = rng.random((500, 3)).astype(np.float32)
    np.testing.assert_allclose(S.hsv2rgb(S.rgb2hsv(rgb)), rgb, atol=2e-6)
