"""numpy front-end of raymarch_oracle.c (CPU restatement of raymarching/src/raymarching.cu).  Test infrastructure.

Function names and argument order mirror the reference's Python wrappers (raymarching/raymarching.py) so parity
tests read like the reference's call sites.  All arrays are numpy (float32 / int32 / uint8).
"""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def morton3D(coords):
    coords = _i32(coords).reshape(-1, 3)
    out = np.empty(coords.shape[0], np.int32)
    lib().oracle_morton3D(_p(coords, C.c_int32), C.c_uint32(coords.shape[0]), _p(out, C.c_int32))
    return out


def morton3D_invert(indices):
    indices = _i32(indices).reshape(-1)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().oracle_morton3D_invert(_p(indices, C.c_int32), C.c_uint32(indices.shape[0]), _p(out, C.c_int32))
    return out


def packbits(grid, thresh):
    grid = _f32(grid).reshape(-1)
    n = grid.shape[0] // 8
    out = np.empty(n, np.uint8)
    lib().oracle_packbits(_p(grid, C.c_float), C.c_uint32(n), C.c_float(thresh), _p(out, C.c_uint8))
    return out


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o = _f32(rays_o).reshape(-1, 3)
    rays_d = _f32(rays_d).reshape(-1, 3)
    aabb = _f32(aabb)
    n = rays_o.shape[0]
    nears = np.empty(n, np.float32)
    fars = np.empty(n, np.float32)
    lib().oracle_near_far_from_aabb(_p(rays_o, C.c_float), _p(rays_d, C.c_float), _p(aabb, C.c_float), C.c_uint32(n),
                                    C.c_float(min_near), _p(nears, C.c_float), _p(fars, C.c_float))
    return nears, fars


def march_rays_train(rays_o, rays_d, bound, bitfield, cascade, H, nears, fars, noises=None, dt_gamma=0.0, max_steps=1024, M=None):
    """Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3], counter [2] (ray-ordered packing)."""
    rays_o = _f32(rays_o).reshape(-1, 3)
    rays_d = _f32(rays_d).reshape(-1, 3)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    nears, fars = _f32(nears), _f32(fars)
    n = rays_o.shape[0]
    if noises is None:
        noises = np.zeros(n, np.float32)
    noises = _f32(noises)
    if M is None:
        M = n * max_steps
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    rays = np.zeros((n, 3), np.int32)
    counter = np.zeros(2, np.int32)
    lib().oracle_march_rays_train(_p(rays_o, C.c_float), _p(rays_d, C.c_float), _p(bitfield, C.c_uint8), C.c_float(bound),
                                  C.c_float(dt_gamma), C.c_uint32(max_steps), C.c_uint32(n), C.c_uint32(cascade), C.c_uint32(H),
                                  C.c_uint32(M), _p(nears, C.c_float), _p(fars, C.c_float), _p(xyzs, C.c_float), _p(dirs, C.c_float),
                                  _p(deltas, C.c_float), _p(rays, C.c_int32), _p(counter, C.c_int32), _p(noises, C.c_float))
    return xyzs, dirs, deltas, rays, counter


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, cascade, H, nears, fars, align=-1,
               noises=None, dt_gamma=0.0, max_steps=1024):
    rays_o = _f32(rays_o).reshape(-1, 3)
    rays_d = _f32(rays_d).reshape(-1, 3)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    rays_alive, rays_t = _i32(rays_alive), _f32(rays_t)
    nears, fars = _f32(nears), _f32(fars)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)  # raymarching.py:331-332
    if noises is None:
        noises = np.zeros(n_alive, np.float32)
    noises = _f32(noises)
    xyzs = np.zeros((M, 3), np.float32)
    dirs = np.zeros((M, 3), np.float32)
    deltas = np.zeros((M, 2), np.float32)
    lib().oracle_march_rays(C.c_uint32(n_alive), C.c_uint32(n_step), _p(rays_alive, C.c_int32), _p(rays_t, C.c_float),
                            _p(rays_o, C.c_float), _p(rays_d, C.c_float), C.c_float(bound), C.c_float(dt_gamma), C.c_uint32(max_steps),
                            C.c_uint32(cascade), C.c_uint32(H), _p(bitfield, C.c_uint8), _p(nears, C.c_float), _p(fars, C.c_float),
                            _p(xyzs, C.c_float), _p(dirs, C.c_float), _p(deltas, C.c_float), _p(noises, C.c_float))
    return xyzs, dirs, deltas


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws = np.empty(N, np.float32)
    depth = np.empty(N, np.float32)
    image = np.empty((N, 3), np.float32)
    lib().oracle_composite_rays_train_forward(_p(sigmas, C.c_float), _p(rgbs, C.c_float), _p(deltas, C.c_float), _p(rays, C.c_int32),
                                              C.c_uint32(M), C.c_uint32(N), C.c_float(T_thresh), _p(ws, C.c_float), _p(depth, C.c_float),
                                              _p(image, C.c_float))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    grad_ws, grad_image, weights_sum, image = _f32(grad_ws), _f32(grad_image), _f32(weights_sum), _f32(image)
    M, N = sigmas.shape[0], rays.shape[0]
    gs = np.zeros(M, np.float32)
    gc = np.zeros((M, 3), np.float32)
    lib().oracle_composite_rays_train_backward(_p(grad_ws, C.c_float), _p(grad_image, C.c_float), _p(sigmas, C.c_float),
                                               _p(rgbs, C.c_float), _p(deltas, C.c_float), _p(rays, C.c_int32), _p(weights_sum, C.c_float),
                                               _p(image, C.c_float), C.c_uint32(M), C.c_uint32(N), C.c_float(T_thresh), _p(gs, C.c_float),
                                               _p(gc, C.c_float))
    return gs, gc


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """In place on rays_alive, rays_t, weights_sum, depth, image (numpy arrays of the right dtype)."""
    assert rays_alive.dtype == np.int32 and rays_t.dtype == np.float32
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    lib().oracle_composite_rays(C.c_uint32(n_alive), C.c_uint32(n_step), C.c_float(T_thresh), _p(rays_alive, C.c_int32),
                                _p(rays_t, C.c_float), _p(sigmas, C.c_float), _p(rgbs, C.c_float), _p(deltas, C.c_float),
                                _p(weights_sum, C.c_float), _p(depth, C.c_float), _p(image, C.c_float))
