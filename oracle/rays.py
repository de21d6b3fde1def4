"""TEST ORACLE (numpy) — get_rays for sampled pixels + ground-truth gather, restating the reference:

    get_rays            nerf/utils.py:54-137  (N > 0, patch_size == 1, error_map None branch :95-97, gathers :114-115,
                                               directions :123-128, origin :130-131)
    GT gather + blend   dnerf/provider.py:340-343 (torch.gather of the flattened image), dnerf/utils.py:61-66 (rgb * a + bg * (1 - a))

Pinned against tests/golden/rays.npz, which tests/golden/make_rays_golden.py produced by executing the reference's OWN
`get_rays` source (extracted from /root/reference/nerf/utils.py; its module cannot be imported here: lpips / tensorboardX /
torchmetrics are absent).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.
"""
import numpy as np


def get_rays(pose, intrinsics, H, W, inds):
    """pose [4,4] cam2world, intrinsics (fx, fy, cx, cy), inds [N] flat pixel indices (h * W + w) -> rays_o, rays_d [N,3] float32.
    fp32 arithmetic in the reference's operation order (the scalar divisions are multiplications by the fp32 reciprocal, as torch
    evaluates tensor / python-scalar)."""
    f32 = np.float32
    fx, fy, cx, cy = [f32(v) for v in intrinsics]
    inds = np.asarray(inds, dtype=np.int64)
    i = (inds % W).astype(f32) + f32(0.5)          # :71-72 (meshgrid of linspace(0, W-1, W), transposed, + 0.5) gathered at inds
    j = (inds // W).astype(f32) + f32(0.5)
    zs = np.ones_like(i)
    xs = (i - cx) * (f32(1.0) / fx) * zs            # :123-125
    ys = (j - cy) * (f32(1.0) / fy) * zs
    d = np.stack([xs, ys, zs], -1)
    d = d / np.sqrt((d * d).sum(-1, keepdims=True, dtype=f32)).astype(f32)   # :127
    R = np.asarray(pose, dtype=f32)[:3, :3]
    rays_d = (d @ R.T).astype(f32)                   # :128
    rays_o = np.broadcast_to(np.asarray(pose, dtype=f32)[:3, 3], rays_d.shape).copy()  # :130-131
    return rays_o, rays_d


def gather_gt(image, inds, bg=None):
    """image [H*W, C] float32 (C = 3 or 4) -> gt_rgb [N,3]: the gathered pixels, alpha-blended on bg [N,3] (None = white) when C = 4."""
    px = np.asarray(image, dtype=np.float32)[np.asarray(inds, dtype=np.int64)]
    if px.shape[-1] == 3:
        return px
    a = px[:, 3:4]
    b = np.ones((px.shape[0], 3), np.float32) if bg is None else np.asarray(bg, dtype=np.float32)
    return (px[:, :3] * a + b * (np.float32(1.0) - a)).astype(np.float32)
