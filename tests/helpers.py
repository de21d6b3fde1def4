"""Shared builders of seeded test inputs (numpy on CPU; the GPU tests upload them)."""
import numpy as np
import torch

from seald_nerf_b200 import synthetic as syn


def scene_bitfield(time_idx=20, H=128, cascade=1, thresh=10.0, time_size=64):
    """Occupancy bitfield [cascade*H^3/8] uint8 of the synthetic figure at one time stamp (+ the density grid)."""
    t = (time_idx + 0.5) / time_size
    coords = syn._morton_coords(H, "cpu").float()
    grids = []
    for cas in range(cascade):
        bound = min(2 ** cas, 2 ** (cascade - 1))
        xyz = (2 * coords / (H - 1) - 1) * (bound - bound / H)
        grids.append(syn.density(xyz, t))
    grid = torch.stack(grids, 0)  # [cascade, H^3]
    bits = syn.pack_bitfield_torch(grid.reshape(1, -1), thresh)[0]
    return bits.numpy(), grid.numpy()


def camera_rays(n, seed=0, radius=3.2, H=800, W=800, pose_seed=0, center_crop=None):
    """n random pixels of one orbit camera -> rays_o, rays_d float32 numpy [n,3]."""
    pose = syn.orbit_poses(1, seed=pose_seed, radius=radius)[0]
    g = torch.Generator().manual_seed(seed)
    if center_crop:
        ii = torch.randint(W // 2 - center_crop, W // 2 + center_crop, (n,), generator=g)
        jj = torch.randint(H // 2 - center_crop, H // 2 + center_crop, (n,), generator=g)
        inds = jj * W + ii
    else:
        inds = torch.randint(0, H * W, (n,), generator=g)
    ro, rd = syn.get_rays(pose, syn.intrinsics(H, W), H, W, inds)
    return ro.numpy().astype(np.float32), rd.numpy().astype(np.float32)


def seal_mapper_from_dict(mp):
    """oracle.seal mapper dict (numpy tensors) -> seald_nerf_b200.SealNeRF.seal_utils mapper object."""
    from seald_nerf_b200.SealNeRF import seal_utils as su
    cls = {"bbox": su.SealBBoxMapper, "brush": su.SealBrushMapper, "anchor": su.SealAnchorMapper}[mp["type"]]
    data = {k: v for k, v in mp.items() if k not in ("type", "map_triangles", "map_test_dir")}
    return cls.from_tensors(data, mp["map_triangles"], mp.get("map_test_dir"))
