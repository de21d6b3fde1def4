"""Drop-in `raymarching` module: same functions, positional arguments and defaults as the reference's
raymarching/raymarching.py (:22 near_far_from_aabb, :55 sph_from_ray, :85 morton3D, :108 morton3D_invert,
:132 packbits, :164 march_rays_train, :241 composite_rays_train, :300 march_rays, :354 composite_rays), calling
libseald_b200.so instead of the pybind11 `_raymarching` backend.  Inputs are cast to fp32 under autocast like the
reference's `custom_fwd(cast_inputs=torch.float32)`.
"""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from .. import _lib
from .._lib import ptr


def _cuda(t):
    return t if t.is_cuda else t.cuda()


class _near_far_from_aabb(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        aabb = _cuda(aabb).contiguous()
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _lib.call("seald_near_far_from_aabb", ptr(rays_o), ptr(rays_d), ptr(aabb), N, float(min_near), ptr(nears), ptr(fars),
                  _lib.stream())
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, radius):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _lib.call("seald_sph_from_ray", ptr(rays_o), ptr(rays_d), float(radius), N, ptr(coords), _lib.stream())
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        coords = _cuda(coords).int().contiguous()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        _lib.call("seald_morton3D", ptr(coords), N, ptr(indices), _lib.stream())
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        indices = _cuda(indices).int().contiguous()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        _lib.call("seald_morton3D_invert", ptr(indices), N, ptr(coords), _lib.stream())
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, grid, thresh, bitfield=None):
        grid = _cuda(grid).contiguous()
        C = grid.shape[0]
        H3 = grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        _lib.call("seald_packbits", ptr(grid), N, float(thresh), ptr(bitfield), _lib.stream())
        return bitfield


packbits = _packbits.apply


class _march_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False,
                align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        density_bitfield = _cuda(density_bitfield).contiguous()

        N = rays_o.shape[0]
        M = N * max_steps
        # running estimate of the sample count from earlier steps (reference: raymarching.py:200-203)
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count

        dev, dt = rays_o.device, rays_o.dtype
        xyzs = torch.zeros(M, 3, dtype=dt, device=dev)
        dirs = torch.zeros(M, 3, dtype=dt, device=dev)
        deltas = torch.zeros(M, 2, dtype=dt, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=dt, device=dev) if perturb else torch.zeros(N, dtype=dt, device=dev)

        _lib.call("seald_march_rays_train", ptr(rays_o), ptr(rays_d), ptr(density_bitfield), float(bound), float(dt_gamma),
                  int(max_steps), N, int(C), int(H), M, ptr(nears.contiguous()), ptr(fars.contiguous()), None, 0.0, None, None,
                  ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(step_counter), ptr(noises), None, _lib.stream())

        # first epochs only: trim to the used length (host sync, as in the reference raymarching.py:223-231)
        if force_all_rays or mean_count <= 0:
            m = step_counter[0].item()
            if align > 0:
                m += align - m % align
            xyzs = xyzs[:m]
            dirs = dirs[:m]
            deltas = deltas[:m]
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        deltas = deltas.contiguous()
        M = sigmas.shape[0]
        N = rays.shape[0]
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        _lib.call("seald_composite_rays_train_forward", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), M, N, float(T_thresh),
                  ptr(weights_sum), ptr(depth), ptr(image), _lib.stream())
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is ignored, like the reference (raymarching.py:274)
        grad_weights_sum = grad_weights_sum.contiguous()
        grad_image = grad_image.contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        _lib.call("seald_composite_rays_train_backward", ptr(grad_weights_sum), ptr(grad_image), ptr(sigmas), ptr(rgbs), ptr(deltas),
                  ptr(rays), ptr(weights_sum), ptr(image), M, N, float(T_thresh), ptr(grad_sigmas), ptr(grad_rgbs), _lib.stream())
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply


class _march_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
                perturb=False, dt_gamma=0, max_steps=1024):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        dev, dt = rays_o.device, rays_o.dtype
        xyzs = torch.zeros(M, 3, dtype=dt, device=dev)
        dirs = torch.zeros(M, 3, dtype=dt, device=dev)
        deltas = torch.zeros(M, 2, dtype=dt, device=dev)
        noises = torch.rand(n_alive, dtype=dt, device=dev) if perturb else torch.zeros(n_alive, dtype=dt, device=dev)
        _lib.call("seald_march_rays", int(n_alive), int(n_step), ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), float(bound),
                  float(dt_gamma), int(max_steps), int(C), int(H), ptr(density_bitfield.contiguous()), ptr(near), ptr(far), ptr(xyzs),
                  ptr(dirs), ptr(deltas), ptr(noises), None, None, None, _lib.stream())
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        _lib.call("seald_composite_rays", int(n_alive), int(n_step), float(T_thresh), ptr(rays_alive), ptr(rays_t),
                  ptr(sigmas.contiguous()), ptr(rgbs.contiguous()), ptr(deltas.contiguous()), ptr(weights_sum), ptr(depth), ptr(image),
                  None, None, _lib.stream())
        return tuple()


composite_rays = _composite_rays.apply


def march_rays_seal(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, mapper, align=-1,
                    perturb=False, dt_gamma=0, max_steps=1024):
    """`march_rays` + `mapper.map_to_origin` in one kernel (SealDNeRF/renderer.py:245-253): returns the MAPPED xyzs / dirs,
    deltas and the map mask (bool [M]).  bbox and brush mappers (`mapper.fusable`)."""
    import ctypes as C_
    rays_o = _cuda(rays_o).float().contiguous().view(-1, 3)
    rays_d = _cuda(rays_d).float().contiguous().view(-1, 3)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    dev, dt = rays_o.device, rays_o.dtype
    xyzs = torch.zeros(M, 3, dtype=dt, device=dev)
    dirs = torch.zeros(M, 3, dtype=dt, device=dev)
    deltas = torch.zeros(M, 2, dtype=dt, device=dev)
    mask = torch.zeros(M, dtype=torch.bool, device=dev)
    noises = torch.rand(n_alive, dtype=dt, device=dev) if perturb else torch.zeros(n_alive, dtype=dt, device=dev)
    _lib.call("seald_march_rays_seal", int(n_alive), int(n_step), ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d), float(bound),
              float(dt_gamma), int(max_steps), int(C), int(H), ptr(density_bitfield.contiguous()), ptr(near), ptr(far), ptr(xyzs),
              ptr(dirs), ptr(deltas), ptr(noises), None, None, C_.byref(mapper.descriptor(dev)), ptr(mask), None, _lib.stream())
    return xyzs, dirs, deltas, mask


def march_rays_train_seal(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, mapper, step_counter=None, mean_count=-1,
                          perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
    """`march_rays_train` + `mapper.map_to_origin` in one kernel (SealDNeRF/renderer.py:150-158): returns the MAPPED xyzs /
    dirs, deltas, rays and the map mask (bool [M]).  No gradient flows through the march (as in the reference)."""
    import ctypes as C_
    rays_o = _cuda(rays_o).float().contiguous().view(-1, 3)
    rays_d = _cuda(rays_d).float().contiguous().view(-1, 3)
    density_bitfield = _cuda(density_bitfield).contiguous()
    N = rays_o.shape[0]
    M = N * max_steps
    if not force_all_rays and mean_count > 0:
        if align > 0:
            mean_count += align - mean_count % align
        M = mean_count
    dev, dt = rays_o.device, rays_o.dtype
    xyzs = torch.zeros(M, 3, dtype=dt, device=dev)
    dirs = torch.zeros(M, 3, dtype=dt, device=dev)
    deltas = torch.zeros(M, 2, dtype=dt, device=dev)
    mask = torch.zeros(M, dtype=torch.bool, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
    noises = torch.rand(N, dtype=dt, device=dev) if perturb else torch.zeros(N, dtype=dt, device=dev)
    _lib.call("seald_march_rays_train_seal", ptr(rays_o), ptr(rays_d), ptr(density_bitfield), float(bound), float(dt_gamma), int(max_steps), N,
              int(C), int(H), M, ptr(nears.contiguous()), ptr(fars.contiguous()), None, 0.0, None, None, ptr(xyzs), ptr(dirs), ptr(deltas),
              ptr(rays), ptr(step_counter), ptr(noises), C_.byref(mapper.descriptor(dev)), ptr(mask), None, _lib.stream())
    if force_all_rays or mean_count <= 0:
        m = step_counter[0].item()
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas, mask = xyzs[:m], dirs[:m], deltas[:m], mask[:m]
    return xyzs, dirs, deltas, rays, mask


def occupancy_aabb(density_bitfield, C, H, bound, guard_cells=2, out=None):
    """World-space box [6] of the occupied cells of one bitfield frame, grown by `guard_cells` cells (seald_occupancy_aabb):
    the optional `occ_aabb6` guard of the march kernels (rays that cannot meet an occupied cell skip the walk)."""
    dev = density_bitfield.device
    if out is None:
        out = torch.empty(6, dtype=torch.float32, device=dev)
    scratch = torch.empty(6 * int(C), dtype=torch.int32, device=dev)
    _lib.call("seald_occupancy_aabb", ptr(density_bitfield.contiguous()), int(C), int(H), float(bound), int(guard_cells), ptr(scratch), ptr(out),
              _lib.stream())
    return out


def compact_alive(rays_alive, n_alive=None, n_alive_dev=None):
    """Device-side, order-preserving replacement of `rays_alive[rays_alive >= 0]` (dnerf/renderer.py:372).
    Returns (compacted int32 [n_alive], count int32 [1] on device) without a host sync."""
    if n_alive is None:
        n_alive = rays_alive.shape[0]
    out = torch.empty(n_alive, dtype=torch.int32, device=rays_alive.device)
    n_out = torch.empty(1, dtype=torch.int32, device=rays_alive.device)
    scratch = torch.empty((n_alive + 1023) // 1024 + 1, dtype=torch.int32, device=rays_alive.device)
    _lib.call("seald_compact_alive", ptr(rays_alive), int(n_alive), ptr(n_alive_dev), ptr(out), ptr(n_out), ptr(scratch),
              _lib.stream())
    return out, n_out
