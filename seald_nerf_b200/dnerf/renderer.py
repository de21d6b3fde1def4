"""Drop-in `dnerf.renderer.NeRFRenderer` (reference: dnerf/renderer.py:61-590): the `cuda_ray=True` render/train path.

`run_cuda`, `update_extra_state`, `mark_untrained_grid`, `reset_extra_state` and `render` keep the reference's
signatures, buffers (`density_grid [64, cascade, 128^3]`, `density_bitfield`, `step_counter [16, 2]`, `aabb_*`, `times`)
and return values.  The non-CUDA sampling path (`run`, cuda_ray=False) is the reference's CPU baseline and is
restated only in oracle/render.py; here it raises.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import raymarching


def custom_meshgrid(*args):
    return torch.meshgrid(*args, indexing="ij")


class NeRFRenderer(nn.Module):
    def __init__(self, bound=1, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1):
        super().__init__()
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.time_size = 64
        self.grid_size = 128
        self.density_scale = density_scale
        self.min_near = min_near
        self.density_thresh = density_thresh
        self.bg_radius = bg_radius

        aabb_train = torch.FloatTensor([-bound, -bound, -bound, bound, bound, bound])
        aabb_infer = aabb_train.clone()
        self.register_buffer("aabb_train", aabb_train)
        self.register_buffer("aabb_infer", aabb_infer)

        self.cuda_ray = cuda_ray
        if cuda_ray:
            density_grid = torch.zeros(self.time_size, self.cascade, self.grid_size ** 3)
            density_bitfield = torch.zeros(self.time_size, self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8)
            self.register_buffer("density_grid", density_grid)
            self.register_buffer("density_bitfield", density_bitfield)
            self.mean_density = 0
            self.iter_density = 0
            times = ((torch.arange(self.time_size, dtype=torch.float32) + 0.5) / self.time_size).view(-1, 1, 1)
            self.register_buffer("times", times)
            step_counter = torch.zeros(16, 2, dtype=torch.int32)
            self.register_buffer("step_counter", step_counter)
            self.mean_count = 0
            self.local_step = 0

    def forward(self, x, d, t):
        raise NotImplementedError()

    def density(self, x, t):
        raise NotImplementedError()

    def color(self, x, d, t, mask=None, **kwargs):
        raise NotImplementedError()

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    def run(self, rays_o, rays_d, time, num_steps=128, upsample_steps=128, bg_color=None, perturb=False, **kwargs):
        raise NotImplementedError("the cuda_ray=False sampling path is the reference's CPU baseline (oracle/render.py); "
                                  "seald_b200 implements the cuda_ray=True hot path only")

    # ---- Seal proxy mapping (SealDNeRF/renderer.py:156-158, 250-253, 271-272); None on a plain D-NeRF model ----------
    seal_mapper = None

    def _march_train(self, rays_o, rays_d, t, nears, fars, counter, perturb, force_all_rays, dt_gamma, max_steps):
        """march_rays_train (+ proxy mapping of the samples when a seal mapper is set) -> xyzs, dirs, deltas, rays, mask"""
        mp = self.seal_mapper
        args = (rays_o, rays_d, self.bound, self.density_bitfield[t], self.cascade, self.grid_size, nears, fars)
        if mp is not None and mp.fusable:
            return raymarching.march_rays_train_seal(*args, mp, counter, self.mean_count, perturb, 128, force_all_rays, dt_gamma, max_steps)
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(*args, counter, self.mean_count, perturb, 128, force_all_rays, dt_gamma,
                                                                max_steps)
        mask = None
        if mp is not None:
            xyzs, dirs, mask = mp.map_to_origin(xyzs.view(-1, 3), dirs.view(-1, 3))
        return xyzs, dirs, deltas, rays, mask

    def _march_infer(self, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bitfield, nears, fars, perturb, dt_gamma, max_steps):
        mp = self.seal_mapper
        args = (n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound, bitfield, self.cascade, self.grid_size, nears, fars)
        if mp is not None and mp.fusable:
            return raymarching.march_rays_seal(*args, mp, 128, perturb, dt_gamma, max_steps)
        xyzs, dirs, deltas = raymarching.march_rays(*args, 128, perturb, dt_gamma, max_steps)
        mask = None
        if mp is not None:
            xyzs, dirs, mask = mp.map_to_origin(xyzs.view(-1, 3), dirs.view(-1, 3))
        return xyzs, dirs, deltas, mask

    def _frame_index(self, time):
        return torch.floor(time[0][0] * self.time_size).clamp(min=0, max=self.time_size - 1).long()

    def run_cuda(self, rays_o, rays_d, time, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024,
                 T_thresh=None, normalize_depth=True, **kwargs):
        # rays_o, rays_d: [B, N, 3] (B == 1); time: [B, 1]  ->  image [B, N, 3], depth [B, N]
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device

        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer, self.min_near)
        if bg_color is None:
            bg_color = 1

        t = self._frame_index(time)
        results = {}

        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1

            # (the reference's training branch maps the samples but leaves map_color commented out, SealDNeRF/renderer.py:180-182)
            xyzs, dirs, deltas, rays, _ = self._march_train(rays_o, rays_d, t, nears, fars, counter, perturb, force_all_rays, dt_gamma, max_steps)
            sigmas, rgbs, deform = self(xyzs, dirs, time)
            sigmas = self.density_scale * sigmas

            if T_thresh is None:
                weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays)
            else:
                weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)
            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
            if normalize_depth:
                depth = torch.clamp(depth - nears, min=0) / (fars - nears)
            image = image.view(*prefix, 3)
            depth = depth.view(*prefix)
            results["deform"] = deform
            results["weights_sum"] = weights_sum
        else:
            dtype = torch.float32
            weights_sum = torch.zeros(N, dtype=dtype, device=device)
            depth = torch.zeros(N, dtype=dtype, device=device)
            image = torch.zeros(N, 3, dtype=dtype, device=device)

            n_alive = N
            rays_alive = torch.arange(n_alive, dtype=torch.int32, device=device)
            rays_t = nears.clone()
            bitfield = self.density_bitfield[t]

            step = 0
            while step < max_steps:
                n_alive = rays_alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas, mask = self._march_infer(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bitfield, nears, fars,
                                                             perturb if step == 0 else False, dt_gamma, max_steps)
                sigmas, rgbs, _ = self(xyzs, dirs, time)
                sigmas = self.density_scale * sigmas
                if mask is not None and self.seal_mapper.has_color_map:
                    # rgbs[mask] = map_color(xyzs[mask], dirs[mask], rgbs[mask]) without the gathers (SealDNeRF/renderer.py:271-272)
                    rgbs = self.seal_mapper.map_color_masked_(xyzs, mask, rgbs.float().contiguous())
                if T_thresh is None:
                    raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image)
                else:
                    raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh)
                # the reference compacts with a boolean mask (host sync on the new length, renderer.py:372); the length is
                # needed on the host anyway to pick n_step, so read it from the device-side compaction
                compacted, n_out = raymarching.compact_alive(rays_alive)
                rays_alive = compacted[:int(n_out.item())]
                step += n_step

            image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
            if normalize_depth:
                depth = torch.clamp(depth - nears, min=0) / (fars - nears)
            image = image.view(*prefix, 3)
            depth = depth.view(*prefix)
            results["weights_sum"] = weights_sum

        results["depth"] = depth
        results["image"] = image
        return results

    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        # cells never covered by a training camera get density -1 (reference: dnerf/renderer.py:388-451)
        if not self.cuda_ray:
            return
        if isinstance(poses, np.ndarray):
            poses = torch.from_numpy(poses)
        B = poses.shape[0]
        fx, fy, cx, cy = intrinsic
        dev = self.density_bitfield.device
        X = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
        Y = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
        Z = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
        count = torch.zeros_like(self.density_grid[0])
        poses = poses.to(count.device)
        for xs in X:
            for ys in Y:
                for zs in Z:
                    xx, yy, zz = custom_meshgrid(xs, ys, zs)
                    coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                    indices = raymarching.morton3D(coords).long()
                    world_xyzs = (2 * coords.float() / (self.grid_size - 1) - 1).unsqueeze(0)
                    for cas in range(self.cascade):
                        bound = min(2 ** cas, self.bound)
                        half_grid_size = bound / self.grid_size
                        cas_world_xyzs = world_xyzs * (bound - half_grid_size)
                        head = 0
                        while head < B:
                            tail = min(head + S, B)
                            cam_xyzs = cas_world_xyzs - poses[head:tail, :3, 3].unsqueeze(1)
                            cam_xyzs = cam_xyzs @ poses[head:tail, :3, :3]
                            mask_z = cam_xyzs[:, :, 2] > 0
                            mask_x = torch.abs(cam_xyzs[:, :, 0]) < cx / fx * cam_xyzs[:, :, 2] + half_grid_size * 2
                            mask_y = torch.abs(cam_xyzs[:, :, 1]) < cy / fy * cam_xyzs[:, :, 2] + half_grid_size * 2
                            mask = (mask_z & mask_x & mask_y).sum(0).reshape(-1)
                            count[cas, indices] += mask
                            head += S
        self.density_grid[count.unsqueeze(0).expand_as(self.density_grid) == 0] = -1

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        # occupancy-grid refresh (reference: dnerf/renderer.py:453-555): same sampling schedule and RNG call order
        if not self.cuda_ray:
            return
        dev = self.density_bitfield.device
        tmp_grid = -torch.ones_like(self.density_grid)

        if self.iter_density < 16:
            X = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
            Y = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
            Z = torch.arange(self.grid_size, dtype=torch.int32, device=dev).split(S)
            for t, time in enumerate(self.times):
                for xs in X:
                    for ys in Y:
                        for zs in Z:
                            xx, yy, zz = custom_meshgrid(xs, ys, zs)
                            coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                            indices = raymarching.morton3D(coords).long()
                            xyzs = 2 * coords.float() / (self.grid_size - 1) - 1
                            for cas in range(self.cascade):
                                bound = min(2 ** cas, self.bound)
                                half_grid_size = bound / self.grid_size
                                half_time_size = 0.5 / self.time_size
                                cas_xyzs = xyzs * (bound - half_grid_size)
                                cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half_grid_size
                                time_perturb = time + (torch.rand_like(time) * 2 - 1) * half_time_size
                                sigmas = self.density(cas_xyzs, time_perturb)["sigma"].reshape(-1).detach()
                                sigmas *= self.density_scale
                                tmp_grid[t, cas, indices] = sigmas
        elif self.iter_density < 100:
            N = self.grid_size ** 3 // 4
            for t, time in enumerate(self.times):
                for cas in range(self.cascade):
                    coords = torch.randint(0, self.grid_size, (N, 3), device=dev)
                    indices = raymarching.morton3D(coords).long()
                    occ_indices = torch.nonzero(self.density_grid[t, cas] > 0).squeeze(-1)
                    rand_mask = torch.randint(0, occ_indices.shape[0], [N], dtype=torch.long, device=dev)
                    occ_indices = occ_indices[rand_mask]
                    occ_coords = raymarching.morton3D_invert(occ_indices)
                    indices = torch.cat([indices, occ_indices], dim=0)
                    coords = torch.cat([coords, occ_coords], dim=0)
                    xyzs = 2 * coords.float() / (self.grid_size - 1) - 1
                    bound = min(2 ** cas, self.bound)
                    half_grid_size = bound / self.grid_size
                    half_time_size = 0.5 / self.time_size
                    cas_xyzs = xyzs * (bound - half_grid_size)
                    cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half_grid_size
                    time_perturb = time + (torch.rand_like(time) * 2 - 1) * half_time_size
                    sigmas = self.density(cas_xyzs, time_perturb)["sigma"].reshape(-1).detach()
                    sigmas *= self.density_scale
                    tmp_grid[t, cas, indices] = sigmas

        valid_mask = (self.density_grid >= 0) & (tmp_grid >= 0)
        self.density_grid[valid_mask] = torch.maximum(self.density_grid[valid_mask] * decay, tmp_grid[valid_mask])
        self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()
        self.iter_density += 1

        density_thresh = min(self.mean_density, self.density_thresh)
        for t in range(self.time_size):
            raymarching.packbits(self.density_grid[t], density_thresh, self.density_bitfield[t])

        total_step = min(16, self.local_step)
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0

    def render(self, rays_o, rays_d, time, staged=False, max_ray_batch=4096, **kwargs):
        # rays_o, rays_d: [B, N, 3] (B == 1) -> dict(image [B, N, 3], depth [B, N], ...); never staged with cuda_ray
        if not self.cuda_ray:
            return self.run(rays_o, rays_d, time, **kwargs)
        return self.run_cuda(rays_o, rays_d, time, **kwargs)
