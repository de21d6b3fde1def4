"""TEST ORACLE (numpy) — the arithmetic of the occupancy-grid refresh, restating NeRFRenderer.update_extra_state
(dnerf/renderer.py:453-555):

    cell_points   :477-497  sample point of a grid cell: (2 c / (H-1) - 1) * (bound - bound/H) + (u * 2 - 1) * bound/H, fp32, with torch's
                            evaluation of `tensor / python-scalar` as a multiplication by the fp32 reciprocal
    morton3D      raymarching.cu:56-70 (10 bits per axis)
    ema_max       :541-543  grid = max(grid * decay, tmp) where grid >= 0 and tmp >= 0 (tmp = -1 marks cells not sampled this pass)
    packbits      raymarching.cu:281-288  bit i of byte n = grid[8n + i] > thresh

Parity pinned on the reference's own expressions evaluated with torch on CPU (tests/test_oracle_occupancy.py); the reference has no
test or golden vector for this function.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.
"""
import numpy as np

f32 = np.float32


def cell_points(coords, u, H, bound):
    """coords [n,3] int, u [n,3] uniform [0,1) float32 -> xyz [n,3] float32"""
    half = bound / H
    x = f32(2.0) * coords.astype(f32) * (f32(1.0) / f32(H - 1)) - f32(1.0)
    return (x * f32(bound - half) + (u.astype(f32) * f32(2.0) - f32(1.0)) * f32(half)).astype(f32)


def _expand_bits(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3D(coords):
    c = np.asarray(coords)
    return (_expand_bits(c[:, 0]) | (_expand_bits(c[:, 1]) << np.uint32(1)) | (_expand_bits(c[:, 2]) << np.uint32(2))).astype(np.int64)


def ema_max(grid, tmp, decay):
    grid = np.asarray(grid, dtype=f32).copy()
    tmp = np.asarray(tmp, dtype=f32)
    valid = (grid >= 0) & (tmp >= 0)
    grid[valid] = np.maximum(grid[valid] * f32(decay), tmp[valid])
    return grid


def packbits(grid, thresh):
    bits = (np.asarray(grid, dtype=f32).reshape(-1, 8) > f32(thresh)).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(-1).astype(np.uint8)
