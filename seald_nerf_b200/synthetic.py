"""Synthetic "jumpingjacks-shaped" D-NeRF scene: the stand-in for dnerf/provider.py's blender loader.

800x800 pinhole cameras (camera_angle_x = 0.6911 rad) on a sphere of radius 4.0 * scale(0.8) looking at the origin
(orbit poses as dnerf/provider.py:56-90), 200 frames with t = i/199, bound 1.  The scene is an analytic,
time-varying density: a union of five capsules (torso, two arms, two legs) whose limb angles swing with
sin(2*pi*t), sigma = 50 inside and 0 outside, with a position-dependent colour.  Everything is plain torch and
runs on whatever device it is given (data generation is set-up work, never part of a timed region).
"""
import math

import numpy as np
import torch

W_DEFAULT = H_DEFAULT = 800
CAMERA_ANGLE_X = 0.6911112070083618


def intrinsics(H=H_DEFAULT, W=W_DEFAULT, camera_angle_x=CAMERA_ANGLE_X):
    fl = W / (2 * np.tan(camera_angle_x / 2))
    return np.array([fl, fl, W / 2, H / 2], dtype=np.float64)


def orbit_poses(n, device="cpu", radius=3.2, theta_range=(math.pi / 3, 2 * math.pi / 3), phi_range=(0.0, 2 * math.pi), seed=0):
    """Look-at-origin poses [n,4,4] (camera-to-world), same construction as dnerf/provider.py:56-90."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    thetas = torch.rand(n, generator=g) * (theta_range[1] - theta_range[0]) + theta_range[0]
    phis = torch.rand(n, generator=g) * (phi_range[1] - phi_range[0]) + phi_range[0]
    centers = torch.stack([radius * torch.sin(thetas) * torch.sin(phis), radius * torch.cos(thetas),
                           radius * torch.sin(thetas) * torch.cos(phis)], dim=-1)

    def normalize(v):
        return v / (torch.norm(v, dim=-1, keepdim=True) + 1e-10)

    forward = -normalize(centers)
    up = torch.tensor([0.0, -1.0, 0.0]).unsqueeze(0).repeat(n, 1)
    right = normalize(torch.cross(forward, up, dim=-1))
    up = normalize(torch.cross(right, forward, dim=-1))
    poses = torch.eye(4).unsqueeze(0).repeat(n, 1, 1)
    poses[:, :3, :3] = torch.stack((right, up, forward), dim=-1)
    poses[:, :3, 3] = centers
    return poses.to(device)


def get_rays(pose, intr, H, W, inds=None):
    """Rays of one camera (nerf/utils.py:54-137 convention: pixel centres at +0.5, directions normalised).
    pose [4,4]; inds: optional flat pixel indices [N]; returns rays_o, rays_d [N,3]."""
    device = pose.device
    fx, fy, cx, cy = [float(v) for v in intr]
    if inds is None:
        inds = torch.arange(H * W, device=device)
    i = (inds % W).float() + 0.5
    j = torch.div(inds, W, rounding_mode="floor").float() + 0.5
    zs = torch.ones_like(i)
    xs = (i - cx) / fx * zs
    ys = (j - cy) / fy * zs
    directions = torch.stack((xs, ys, zs), dim=-1)
    directions = directions / torch.norm(directions, dim=-1, keepdim=True)
    rays_d = directions @ pose[:3, :3].transpose(-1, -2)
    rays_o = pose[:3, 3].unsqueeze(0).expand_as(rays_d)
    return rays_o.contiguous(), rays_d.contiguous()


def _capsules(t):
    """Five capsules (a, b, radius) at time t (python float)."""
    s = math.sin(2 * math.pi * t)
    arm = math.radians(20 + 50 * (0.5 + 0.5 * s))  # arm elevation from the body axis
    leg = math.radians(5 + 20 * (0.5 + 0.5 * s))
    k = 1.5  # overall size of the figure inside the [-1,1]^3 box
    caps = [((0.0, -0.05 * k, 0.0), (0.0, 0.35 * k, 0.0), 0.11 * k)]  # torso (y up)
    for sx in (-1.0, 1.0):
        sh = (0.12 * k * sx, 0.30 * k, 0.0)
        caps.append((sh, (sh[0] + sx * 0.42 * k * math.sin(arm), sh[1] - 0.42 * k * math.cos(arm), 0.0), 0.05 * k))
        hp = (0.07 * k * sx, -0.05 * k, 0.0)
        caps.append((hp, (hp[0] + sx * 0.55 * k * math.sin(leg), hp[1] - 0.55 * k * math.cos(leg), 0.0), 0.06 * k))
    return caps


def density(xyz, t, sigma_in=50.0):
    """Analytic density at points xyz [N,3] (torch) and time t (float): sigma_in inside any capsule else 0."""
    out = torch.zeros(xyz.shape[0], device=xyz.device, dtype=torch.float32)
    for a, b, r in _capsules(float(t)):
        a = torch.tensor(a, device=xyz.device)
        b = torch.tensor(b, device=xyz.device)
        ab = b - a
        h = ((xyz - a) @ ab / (ab @ ab)).clamp(0, 1)
        d = torch.norm(xyz - a - h.unsqueeze(-1) * ab, dim=-1)
        out = torch.where(d < r, torch.full_like(out, sigma_in), out)
    return out


def color(xyz):
    return (0.5 + 0.5 * torch.sin(xyz * torch.tensor([7.0, 5.0, 9.0], device=xyz.device) + torch.tensor([0.0, 1.0, 2.0], device=xyz.device))).float()


def _morton_coords(H, device):
    """coords [H^3, 3] such that coords[m] is the cell whose Morton code is m (raymarching.cu:73-81)."""
    m = torch.arange(H ** 3, device=device, dtype=torch.int64)

    def compact(x):
        x = x & 0x49249249
        x = (x | (x >> 2)) & 0xC30C30C3
        x = (x | (x >> 4)) & 0x0F00F00F
        x = (x | (x >> 8)) & 0xFF0000FF
        x = (x | (x >> 16)) & 0x0000FFFF
        return x

    return torch.stack([compact(m), compact(m >> 1), compact(m >> 2)], dim=-1)


def make_density_grid(time_size=64, H=128, bound=1.0, device="cpu", sigma_in=50.0):
    """density_grid [T, 1, H^3] fp32 in Morton order, sampled at cell centres at the grid's time stamps
    (dnerf/renderer.py:92, :478-485 coordinate convention: xyz = 2*c/(H-1) - 1 scaled by (bound - bound/H))."""
    coords = _morton_coords(H, device).float()
    xyz = (2 * coords / (H - 1) - 1) * (bound - bound / H)
    grid = torch.empty(time_size, 1, H ** 3, device=device)
    for ti in range(time_size):
        grid[ti, 0] = density(xyz, (ti + 0.5) / time_size, sigma_in)
    return grid


def pack_bitfield_torch(density_grid, thresh):
    """Pure-torch bit packing (bit i of byte n = grid[8n+i] > thresh); used for set-up on CPU or GPU."""
    T = density_grid.shape[0]
    occ = (density_grid.reshape(T, -1, 8) > thresh).to(torch.uint8)
    weights = (2 ** torch.arange(8, device=density_grid.device)).to(torch.uint8)
    return (occ * weights).sum(-1).to(torch.uint8)


def render_gt(rays_o, rays_d, t, n_samples=256, bound=1.0, min_near=0.2):
    """Reference image of the analytic scene with fixed-step quadrature (set-up only)."""
    device = rays_o.device
    rd = 1.0 / rays_d
    t0 = (-bound - rays_o) * rd
    t1 = (bound - rays_o) * rd
    near = torch.minimum(t0, t1).amax(-1).clamp(min=min_near)
    far = torch.maximum(t0, t1).amin(-1)
    hit = far > near
    z = near.unsqueeze(-1) + (far - near).clamp(min=0).unsqueeze(-1) * torch.linspace(0, 1, n_samples, device=device)
    pts = rays_o.unsqueeze(1) + rays_d.unsqueeze(1) * z.unsqueeze(-1)
    sig = density(pts.reshape(-1, 3), t).reshape(z.shape)
    col = color(pts.reshape(-1, 3)).reshape(*z.shape, 3)
    dz = ((far - near).clamp(min=0) / n_samples).unsqueeze(-1)
    alpha = 1 - torch.exp(-sig * dz)
    T = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * T * hit.unsqueeze(-1)
    rgb = (w.unsqueeze(-1) * col).sum(1)
    return rgb, w.sum(1)
