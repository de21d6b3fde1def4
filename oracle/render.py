"""The reference's pure-PyTorch render path on CPU (BASELINE.json configs[0]) — a port used ONLY as the timed CPU baseline
(`bench.py` cpu_baseline / `--impl reference`) and as a loose cross-check.  TEST/BENCH INFRASTRUCTURE ONLY.

Follows: NeRFRenderer.run with cuda_ray=False (dnerf/renderer.py:129-258: linspace sampling between the AABB
near/far, alpha compositing with cumprod, masked colour query, white background), the D-NeRF field of
dnerf/network.py:123-257 (density() then color(mask)), with the torch FreqEncoder of encoding.py:5-42 standing in for
every encoder (xyz multires 10, time 6, canonical 10, dirs 4) because the as-shipped `frequency`/`sphere_harmonics`/grid
encoders are CUDA-only, and the slab test of raymarching.cu:92-145 in torch.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class TorchFreqEncoder(nn.Module):
    """encoding.py:5-42 (log-sampled bands, include_input, sin & cos)."""

    def __init__(self, input_dim, multires):
        super().__init__()
        self.freq_bands = (2.0 ** torch.linspace(0.0, multires - 1, multires)).tolist()
        self.output_dim = input_dim + input_dim * multires * 2

    def forward(self, x, **kw):
        out = [x]
        for f in self.freq_bands:
            out.append(torch.sin(x * f))
            out.append(torch.cos(x * f))
        return torch.cat(out, dim=-1)


def near_far_from_aabb(rays_o, rays_d, aabb, min_near):
    rd = 1.0 / rays_d
    t0 = (aabb[:3] - rays_o) * rd
    t1 = (aabb[3:] - rays_o) * rd
    near = torch.minimum(t0, t1).amax(-1)
    far = torch.maximum(t0, t1).amin(-1)
    miss = near > far
    near = near.clamp(min=min_near)
    big = torch.finfo(torch.float32).max
    near = torch.where(miss, torch.full_like(near, big), near)
    far = torch.where(miss, torch.full_like(far, big), far)
    return near, far


class CPUReferenceDNeRF(nn.Module):
    def __init__(self, bound=1.0, min_near=0.2, density_scale=1.0, num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3,
                 hidden_dim_color=64, num_layers_deform=8, hidden_dim_deform=128):
        super().__init__()
        self.bound, self.min_near, self.density_scale = bound, min_near, density_scale
        self.register_buffer("aabb", torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32))
        self.encoder_deform = TorchFreqEncoder(3, 10)
        self.encoder_time = TorchFreqEncoder(1, 6)
        self.encoder = TorchFreqEncoder(3, 10)
        self.encoder_dir = TorchFreqEncoder(3, 4)

        def stack(n, i, h, o):
            return nn.ModuleList([nn.Linear(i if l == 0 else h, o if l == n - 1 else h, bias=False) for l in range(n)])

        self.deform_net = stack(num_layers_deform, self.encoder_deform.output_dim + self.encoder_time.output_dim, hidden_dim_deform, 3)
        self.sigma_net = stack(num_layers, self.encoder.output_dim, hidden_dim, 1 + geo_feat_dim)
        self.color_net = stack(num_layers_color, self.encoder_dir.output_dim + geo_feat_dim, hidden_dim_color, 3)

    @staticmethod
    def _mlp(net, h):
        for l, layer in enumerate(net):
            h = layer(h)
            if l != len(net) - 1:
                h = F.relu(h, inplace=True)
        return h

    def density(self, x, t):
        enc_t = self.encoder_time(t).repeat(x.shape[0], 1)
        deform = self._mlp(self.deform_net, torch.cat([self.encoder_deform(x), enc_t], dim=1))
        if t != 0:
            x = x + deform
        h = self._mlp(self.sigma_net, self.encoder(x))
        return {"sigma": torch.exp(h[..., 0]), "geo_feat": h[..., 1:], "deform": deform}

    def color(self, x, d, mask, geo_feat):
        rgbs = torch.zeros(mask.shape[0], 3, dtype=x.dtype)
        if not mask.any():
            return rgbs
        h = torch.cat([self.encoder_dir(d[mask]), geo_feat[mask]], dim=-1)
        rgbs[mask] = torch.sigmoid(self._mlp(self.color_net, h))
        return rgbs

    @torch.no_grad()
    def run(self, rays_o, rays_d, time, num_steps=128, bg_color=1.0):
        """dnerf/renderer.py:129-258 with upsample_steps=0, perturb=False."""
        rays_o = rays_o.reshape(-1, 3)
        rays_d = rays_d.reshape(-1, 3)
        N = rays_o.shape[0]
        nears, fars = near_far_from_aabb(rays_o, rays_d, self.aabb, self.min_near)
        nears, fars = nears.unsqueeze(-1), fars.unsqueeze(-1)
        z_vals = torch.linspace(0.0, 1.0, num_steps).unsqueeze(0).expand(N, num_steps)
        z_vals = nears + (fars - nears) * z_vals
        sample_dist = (fars - nears) / num_steps
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, self.aabb[:3]), self.aabb[3:])
        dens = self.density(xyzs.reshape(-1, 3), time)
        sigma = dens["sigma"].view(N, num_steps)
        deltas = z_vals[..., 1:] - z_vals[..., :-1]
        deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
        alphas = 1 - torch.exp(-deltas * self.density_scale * sigma)
        alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        mask = weights > 1e-4
        rgbs = self.color(xyzs.reshape(-1, 3), dirs.reshape(-1, 3), mask.reshape(-1), dens["geo_feat"].view(N * num_steps, -1)).view(N, -1, 3)
        weights_sum = weights.sum(dim=-1)
        ori_z = ((z_vals - nears) / (fars - nears)).clamp(0, 1)
        depth = torch.sum(weights * ori_z, dim=-1)
        image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2) + (1 - weights_sum).unsqueeze(-1) * bg_color
        return {"image": image, "depth": depth, "weights_sum": weights_sum}


def time_cpu_render(n_rays=4096, num_steps=128, steps=3, warmup=1, seed=0):
    """Times forward+render of `n_rays` rays on all host cores; returns (rays_per_s, cores, seconds_per_batch)."""
    import os
    import time as _time
    torch.manual_seed(seed)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = CPUReferenceDNeRF()
    g = torch.Generator().manual_seed(seed)
    # rays of an orbit camera at radius 3.2 looking at the origin (same geometry as the GPU workload)
    o = torch.tensor([[0.0, 0.0, 3.2]]).repeat(n_rays, 1)
    d = F.normalize(torch.cat([(torch.rand(n_rays, 2, generator=g) - 0.5) * 0.6, -torch.ones(n_rays, 1)], dim=1), dim=-1)
    t = torch.tensor([[0.5]])
    for _ in range(warmup):
        model.run(o, d, t, num_steps)
    t0 = _time.perf_counter()
    for _ in range(steps):
        model.run(o, d, t, num_steps)
    dt = (_time.perf_counter() - t0) / steps
    return n_rays / dt, cores, dt
