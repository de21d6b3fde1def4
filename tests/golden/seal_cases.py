"""Seeded Seal proxy-mapping cases shared by make_seal_golden.py (reference side), the CPU oracle tests and the GPU parity
tests.  Each case: (mapper dict of numpy tensors, points [P,3], dirs [P,3], colours [P,3])."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def _points(mp, n, seed):
    """Points concentrated around the edit volume (so a good share maps) plus scene-wide ones and zero padding rows."""
    rng = np.random.default_rng(seed)
    b = np.asarray(mp["map_bound"], np.float32).reshape(-1, 2, 3)
    lo, hi = b[:, 0].min(0), b[:, 1].max(0)
    ext = hi - lo
    near = rng.uniform(lo - 0.3 * ext, hi + 0.3 * ext, (n * 3 // 4, 3))
    far = rng.uniform(-1, 1, (n - near.shape[0], 3))
    pts = np.concatenate([near, far]).astype(np.float32)
    pts[-7:] = 0.0           # zero rows (march padding): never mapped
    pts[-9, 1] = 0.0         # one exactly-zero coordinate inside the volume is excluded too (`points.all(1)`)
    pts[-9, [0, 2]] = ((lo + hi) / 2)[[0, 2]]
    dirs = rng.standard_normal((n, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    cols = rng.random((n, 3)).astype(np.float32) * 0.9 + 0.05
    cols[:5] = cols[:5, :1]  # grey colours (delta == 0 branch of rgb2hsv)
    return pts, dirs, cols


def cases():
    from oracle import seal as S
    out = {}
    specs = {
        "bbox_to": S.make_bbox_mapper(),
        "bbox_scale_src": S.make_bbox_mapper(scale=(1.2, 0.8, 1.0), rot_deg=-20.0, translate=(0.1, 0.1, -0.05), map_source=(0.9, 0.9, 0.9),
                                             hsv=(0.05, -0.1, 0.02)),
        "bbox_both_rgb": S.make_bbox_mapper(bound_type="both", rgb=(0.8, 0.2, 0.1), light_offset=0.1),
        "brush_linear": S.make_brush_mapper(mode="linear"),
        "brush_dry_rgb": S.make_brush_mapper(mode="dry", rgb=(1.0, 0.0, 0.0)),
        "brush_image": S.make_brush_mapper(mode="linear", image=True),
        "anchor": S.make_anchor_mapper(scale=(1.0, 1.1, 0.9)),
        "bbox_none": S.make_bbox_mapper(center=(3.0, 3.0, 3.0), translate=(0.0, 0.0, 0.0), rot_deg=0.0),
    }
    for k, (name, mp) in enumerate(specs.items()):
        pts, dirs, cols = _points(mp, 2000, 40 + k)
        if name == "bbox_none":  # nothing falls in the volume: the early-exit path
            pts = np.clip(pts, -1, 1)
        out[name] = (mp, pts, dirs, cols)
    return out
