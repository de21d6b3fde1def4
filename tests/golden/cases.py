"""Seeded inputs of the golden-vector cases (shared by make_golden.py on the GPU box and the CPU oracle tests)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import camera_rays, scene_bitfield  # noqa: E402


def raymarch_case():
    H, cascade, bound = 128, 1, 1.0
    bits, grid = scene_bitfield(time_idx=20, H=H, cascade=cascade)
    ro, rd = camera_rays(512, seed=21, center_crop=220)
    rd[:8] = np.array([0, 0, -1], np.float32)          # axis-parallel rays
    ro[8:24] = np.random.default_rng(3).uniform(-0.4, 0.4, (16, 3)).astype(np.float32)  # rays starting inside the box
    rng = np.random.default_rng(7)
    dens = rng.normal(0, 10, 8 * 4096).astype(np.float32)
    dens[rng.integers(0, dens.size, 500)] = -1
    return dict(H=H, cascade=cascade, bound=bound, min_near=0.2, max_steps=1024, bitfield=bits, rays_o=ro, rays_d=rd,
                aabb=np.array([-1, -1, -1, 1, 1, 1], np.float32), noises=rng.random(512, dtype=np.float32),
                coords=rng.integers(0, 128, (4096, 3)).astype(np.int32), density=dens, thresh=float(dens[11]))


def composite_inputs(M, N, seed=1):
    rng = np.random.default_rng(seed)
    sig = (rng.random(M, dtype=np.float32) * 30).astype(np.float32)
    rgb = rng.random((M, 3), dtype=np.float32)
    gws = rng.random(N, dtype=np.float32)
    gim = rng.random((N, 3), dtype=np.float32)
    return sig, rgb, gws, gim


def grid_cases():
    from oracle import grid as og
    out = {}
    specs = {
        "hash3d": dict(D=3, L=8, C=2, log2_T=12, base=8, desired=512, gridtype=0, align=False, interp=0),
        "tiled4d": dict(D=4, L=6, C=2, log2_T=12, base=4, desired=64, gridtype=1, align=False, interp=0),
        "smooth3d": dict(D=3, L=6, C=4, log2_T=11, base=4, desired=128, gridtype=0, align=True, interp=1),
    }
    for k, (name, s) in enumerate(specs.items()):
        offsets, pls = og.make_offsets(s["D"], s["L"], s["C"], 2.0, s["base"], s["log2_T"], s["desired"], s["align"])
        rng = np.random.default_rng(100 + k)
        B = 384
        x = rng.random((B, s["D"]), dtype=np.float32)
        x[0] = 0.0
        x[1] = 1.0
        x[2, 0] = -1e-6
        x[3] = x[4]
        table = (rng.standard_normal((int(offsets[-1]), s["C"])) * 0.1).astype(np.float32)
        grad = rng.standard_normal((B, s["L"] * s["C"])).astype(np.float32)
        out[name] = dict(x=x, table=table, offsets=offsets, grad=grad, S=float(np.log2(pls)), H=s["base"], L=s["L"], C=s["C"],
                         gridtype=s["gridtype"], align=s["align"], interp=s["interp"])
    return out


def encoder_case():
    rng = np.random.default_rng(11)
    x = (rng.random((300, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    t = rng.random((300, 1), dtype=np.float32)
    d = rng.standard_normal((300, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[0] = 0
    return dict(x=x, t=t, dirs=d, g_x=rng.standard_normal((300, 63)).astype(np.float32), g_t=rng.standard_normal((300, 13)).astype(np.float32))
