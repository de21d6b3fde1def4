"""Drop-in `SealDNeRF.renderer` (reference: SealDNeRF/renderer.py:27-297): the D-NeRF renderer with a Seal proxy mapper.

`SealNeRFTeacherRenderer.run_cuda` keeps the reference's signature and behaviour — `T_thresh` defaults to 1e-4, depth is
NOT normalised (:207, :284), `weights_sum` is returned, mapped samples are queried instead of the marched ones and (in
the inference branch) the colours of mapped samples go through `map_color` — but the march, the mapping, the field and
the compositing run as the fused sm_100a kernels of `dnerf.renderer.NeRFRenderer` (bbox / brush mapping happens INSIDE
the march kernel: `seald_march_rays_seal`).
"""
import torch

from .. import raymarching
from ..dnerf.renderer import NeRFRenderer
from ..SealNeRF.seal_utils import get_seal_mapper


class SealNeRFRenderer(NeRFRenderer):
    def __init__(self, bound=1, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1, **kwargs):
        super().__init__(bound=bound, cuda_ray=cuda_ray, density_scale=density_scale, min_near=min_near, density_thresh=density_thresh,
                         bg_radius=bg_radius)
        self.seal_mapper = None
        self.density_bitfield_origin = None
        self.density_bitfield_hacked = False

    def init_mapper(self, config_dir="", config_dict=None, config_file="seal.json", mapper=None):
        # reference: SealDNeRF/renderer.py:40-74 — also precomputes the occupancy cells inside `force_fill_bound`
        self.seal_mapper = get_seal_mapper(config_dir, config_dict, config_file) if mapper is None else mapper
        bounds = torch.as_tensor(self.seal_mapper.map_data["force_fill_bound"], dtype=torch.float32).clone()
        if bounds.ndim == 2:
            bounds = bounds[None]
        aabb = self.aabb_infer.detach().cpu()
        bounds[:, 0, :] = torch.max(bounds[:, 0, :], aabb[:3])
        bounds[:, 1, :] = torch.min(bounds[:, 1, :], aabb[-3:])
        grid_indices, bitfield_indices = [], []
        dev = self.density_bitfield.device
        for i in range(bounds.shape[0]):
            cmin, cmax = torch.floor(((bounds[i] + self.bound) / self.bound / 2) * self.grid_size)
            X, Y, Z = torch.meshgrid(torch.arange(cmin[0], cmax[0]), torch.arange(cmin[1], cmax[1]), torch.arange(cmin[2], cmax[2]), indexing="ij")
            coords = torch.stack([X, Y, Z], dim=-1).reshape(-1, 3)
            idx = raymarching.morton3D(coords.to(dev)).long()
            grid_indices.append(idx)
            bitfield_indices.append(idx // 8)
        self.force_fill_grid_indices = torch.concat(grid_indices)
        self.force_fill_bitfield_indices = torch.concat(bitfield_indices)

    def update_extra_state(self, decay=0.95, S=128):
        super().update_extra_state(decay, S)
        if self.seal_mapper is not None:
            self.hack_bitfield()

    @torch.no_grad()
    def hack_grids(self):
        self.density_grid[:, self.force_fill_grid_indices] = min(self.mean_density * 1.5, self.density_thresh) + 1e-5

    @torch.no_grad()
    def hack_bitfield(self):
        # the SealD reference only raises the flag here (the force-fill is commented out, SealDNeRF/renderer.py:88-99)
        self.density_bitfield_hacked = True

    @torch.no_grad()
    def restore_bitfield(self):
        for t in range(self.time_size):
            self.density_bitfield[t][self.force_fill_bitfield_indices] = self.density_bitfield_origin
        self.density_bitfield_hacked = False


class SealNeRFTeacherRenderer(SealNeRFRenderer):
    def __init__(self, bound=1, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1, log2_hashmap_size=18,
                 **kwargs):
        super().__init__(bound=bound, cuda_ray=cuda_ray, density_scale=density_scale, min_near=min_near, density_thresh=density_thresh,
                         bg_radius=bg_radius)

    def run_cuda(self, rays_o, rays_d, time, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024, T_thresh=1e-4,
                 **kwargs):
        # reference: SealDNeRF/renderer.py:114-291
        self.time_frame = self._frame_index(time)
        return super().run_cuda(rays_o, rays_d, time, dt_gamma=dt_gamma, bg_color=bg_color, perturb=perturb, force_all_rays=force_all_rays,
                                max_steps=max_steps, T_thresh=T_thresh, normalize_depth=False)


class SealNeRFStudentRenderder(SealNeRFRenderer):
    # (sic: the reference spells it this way, SealDNeRF/renderer.py:294)
    def __init__(self, bound=1, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1, log2_hashmap_size=18,
                 **kwargs):
        super().__init__(bound=bound, cuda_ray=cuda_ray, density_scale=density_scale, min_near=min_near, density_thresh=density_thresh,
                         bg_radius=bg_radius)
