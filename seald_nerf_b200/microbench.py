"""Per-kernel roofline microbenchmarks (BASELINE.json configs[4] and the march / composite / packbits stages).

Every number is device-timed with CUDA events on the launching stream after warm-up; working sets are larger than the
126 MB L2 or an L2 flush (a 256 MiB memset) runs between timed launches, as stated per entry.  `algorithmic bytes` are the
SURVEY.md §8(d) per-unit figures x the units one launch processes.
"""
import json
import os
import sys

import numpy as np
import torch

from . import _lib
from ._lib import ptr, F16, F32


def _time(fn, reps=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def grid_encoder(device, log2_B=22, D=3, gridtype="hash", dtype=torch.float16, coherent=False, reps=20, hbm_gbs=6537.6):
    """GridEncoder fwd / bwd at B = 2^log2_B points, L16 / T2^19 / F2 (configs[4])."""
    from .gridencoder import GridEncoder
    B = 1 << log2_B
    enc = GridEncoder(input_dim=D, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048,
                      gridtype=gridtype).to(device)
    g = torch.Generator(device=device).manual_seed(0)
    x = torch.rand(B, D, device=device, generator=g)
    if coherent:
        # ray-coherent order: 128 consecutive samples lie on one short segment, like marched samples
        seg = torch.rand(B // 128, 1, D, device=device, generator=g)
        dirs = torch.nn.functional.normalize(torch.randn(B // 128, 1, D, device=device, generator=g), dim=-1)
        tt = torch.arange(128, device=device).view(1, 128, 1) * (2 * 3 ** 0.5 / 1024 / 2)
        x = (seg + dirs * tt).clamp(0, 1).reshape(B, D).contiguous()
    table = enc.embeddings.detach().to(dtype).contiguous()
    table.uniform_(-0.1, 0.1)
    out = torch.empty(B, 32, device=device, dtype=dtype)
    grad = torch.randn(B, 32, device=device).to(dtype)
    gtab = torch.zeros(table.shape, device=device, dtype=torch.float32)
    gtab16 = torch.zeros(table.shape, device=device, dtype=dtype)
    gx = torch.empty(B, D, device=device)
    S, H = float(np.log2(enc.per_level_scale)), 16
    dt = F16 if dtype == torch.float16 else F32
    gt = enc.gridtype_id
    s = 2 if dtype == torch.float16 else 4
    st = _lib.stream

    def fwd():
        _lib.call("seald_grid_encode_forward", ptr(x), ptr(table), ptr(enc.offsets), ptr(out), None, B, D, 2, 16, S, H, gt, 0, 0, dt, None, st())

    def bwd32():
        _lib.call("seald_grid_encode_backward", ptr(grad), ptr(x), ptr(table), ptr(enc.offsets), ptr(gtab), None, None, B, D, 2, 16, S, H, gt, 0,
                  0, dt, F32, None, st())

    def bwd16():
        _lib.call("seald_grid_encode_backward", ptr(grad), ptr(x), ptr(table), ptr(enc.offsets), ptr(gtab16), None, None, B, D, 2, 16, S, H, gt,
                  0, 0, dt, dt, None, st())

    def bwd_x():
        _lib.call("seald_grid_encode_backward", ptr(grad), ptr(x), ptr(table), ptr(enc.offsets), ptr(gtab), None, ptr(gx), B, D, 2, 16, S, H, gt,
                  0, 0, dt, F32, None, st())

    res = {}
    nc = 1 << D
    bytes_fwd = B * (4 * D + nc * 16 * 2 * s + 16 * 2 * s)
    bytes_bwd = B * (4 * D + 16 * 2 * s + 2 * nc * 16 * 2 * s)
    bytes_bwd32 = B * (4 * D + 16 * 2 * s + 2 * nc * 16 * 2 * 4)
    for name, fn, nbytes in (("fwd", fwd, bytes_fwd), ("bwd_table_f32", bwd32, bytes_bwd32), ("bwd_table_same_dtype", bwd16, bytes_bwd),
                             ("bwd_table_f32+input_grad", bwd_x, bytes_bwd32 + B * (4 * D + nc * 16 * 2 * s))):
        ms = _time(fn, reps=reps)
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 4), "GB/s": round(gbs, 1), "frac_of_hbm": round(gbs / hbm_gbs, 4),
                     "Gpoints/s": round(B / (ms * 1e-3) / 1e9, 3)}
    return {"B": B, "D": D, "gridtype": gridtype, "dtype": str(dtype).replace("torch.", ""), "order": "coherent" if coherent else "uniform",
            "l2": "inputs+outputs (%d MB) exceed L2; the %.1f MiB table stays L2 resident by design" % ((B * (4 * D + 64)) >> 20, table.numel() * s / 2 ** 20),
            "kernels": res}


def march_composite(device, n_rays=640000, reps=10, hbm_gbs=6537.6):
    """Training march + composite fwd/bwd + packbits on a whole 800x800 frame of the synthetic scene."""
    from . import synthetic as syn
    from . import raymarching as rm
    H = 128
    grid = syn.make_density_grid(64, H, 1.0, device)
    bits = torch.empty(H ** 3 // 8, dtype=torch.uint8, device=device)
    rm.packbits(grid[20], 10.0, bits)
    pose = syn.orbit_poses(1, device, seed=0)[0]
    inds = torch.arange(800 * 800, device=device)[:n_rays]
    ro, rd = syn.get_rays(pose, syn.intrinsics(), 800, 800, inds)
    ro, rd = ro.contiguous(), rd.contiguous()
    N = ro.shape[0]
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=device)
    counter = torch.zeros(2, dtype=torch.int32, device=device)
    nears, fars = rm.near_far_from_aabb(ro, rd, aabb, 0.2)
    rm.march_rays_train(ro, rd, 1.0, bits, 1, H, nears, fars, counter, -1, False, 128, True, 0, 1024)
    m_live = int(counter[0])
    M = (m_live + 127) // 128 * 128
    xyzs = torch.zeros(M, 3, device=device); dirs = torch.zeros(M, 3, device=device); deltas = torch.zeros(M, 2, device=device)
    rays = torch.zeros(N, 3, dtype=torch.int32, device=device)
    noises = torch.zeros(N, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    st = _lib.stream

    occ = rm.occupancy_aabb(bits, 1, H, 1.0, 2)
    use_occ = False

    def march():
        counter.zero_()
        _lib.call("seald_march_rays_train", ptr(ro), ptr(rd), ptr(bits), 1.0, 0.0, 1024, N, 1, H, M, None, None, ptr(aabb), 0.2, ptr(nears),
                  ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(counter), ptr(noises), ptr(occ) if use_occ else None, st())

    sig = torch.rand(M, device=device) * 30
    rgb = torch.rand(M, 3, device=device)
    ws = torch.empty(N, device=device); depth = torch.empty(N, device=device); image = torch.empty(N, 3, device=device)
    gws = torch.rand(N, device=device); gim = torch.rand(N, 3, device=device)
    gs = torch.zeros(M, device=device); gc = torch.zeros(M, 3, device=device)

    def comp_fwd():
        _lib.call("seald_composite_rays_train_forward", ptr(sig), ptr(rgb), ptr(deltas), ptr(rays), M, N, 1e-4, ptr(ws), ptr(depth), ptr(image), st())

    def comp_bwd():
        _lib.call("seald_composite_rays_train_backward", ptr(gws), ptr(gim), ptr(sig), ptr(rgb), ptr(deltas), ptr(rays), ptr(ws), ptr(image), M, N,
                  1e-4, ptr(gs), ptr(gc), st())

    g1 = grid[20].contiguous()

    def pack():
        _lib.call("seald_packbits", ptr(g1), H ** 3 // 8, 10.0, ptr(bits), st())

    gall = grid.reshape(-1).contiguous()
    ball = torch.empty(64 * H ** 3 // 8, dtype=torch.uint8, device=device)

    def pack_all():
        _lib.call("seald_packbits", ptr(gall), 64 * H ** 3 // 8, 10.0, ptr(ball), st())

    out = {"rays": N, "samples": m_live, "l2": "256 MiB flush between timed launches"}
    ms0 = _time(march, reps=reps, flush=flush)
    out["march_train_no_occupancy_guard"] = {"ms": round(ms0, 4)}
    use_occ = True
    march()
    assert int(counter[0]) == m_live, "the occupied-region guard must not change the sample count"
    for name, fn, nbytes in (("march_train", march, 48.0 * N + 32.0 * m_live + 262144.0),
                             ("composite_fwd", comp_fwd, 24.0 * m_live + 32.0 * N),
                             ("composite_bwd", comp_bwd, 40.0 * m_live + 48.0 * N),
                             ("packbits_1frame", pack, 4.125 * H ** 3),
                             ("packbits_64frames", pack_all, 4.125 * 64 * H ** 3)):
        ms = _time(fn, reps=reps, flush=flush)
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 5), "GB/s": round(gbs, 1), "frac_of_hbm": round(gbs / hbm_gbs, 4)}
    return out


def build_scene(device, seed=0, seald=False):
    """Random-init D-NeRF (hashgrid) + analytic occupancy grid of the synthetic jumpingjacks-shaped figure."""
    from . import synthetic as syn
    from . import raymarching
    if seald:
        from .SealDNeRF.network import NeRFNetwork
    else:
        from .dnerf.network import NeRFNetwork
    torch.manual_seed(seed)
    model = NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10).to(device)
    grid = syn.make_density_grid(model.time_size, model.grid_size, 1.0, device)
    model.density_grid.copy_(grid)
    model.mean_density = float(grid.clamp(min=0).mean())
    thresh = min(model.mean_density, model.density_thresh)
    for t in range(model.time_size):
        raymarching.packbits(model.density_grid[t], thresh, model.density_bitfield[t])
    return model


def frame_rays(device, frame=0, H=800, W=800, seed=0):
    from . import synthetic as syn
    pose = syn.orbit_poses(200, device, seed=seed)[frame]
    inds = torch.arange(H * W, device=device)
    ro, rd = syn.get_rays(pose, syn.intrinsics(H, W), H, W, inds)
    return ro.contiguous(), rd.contiguous()


def frame_render(device, model=None, times=(0.0, 0.25, 0.5, 0.75, 1.0), reps=3, rank=0, world_size=1, T_thresh=1e-2):
    """Full 800x800 frame (BASELINE.json configs[2]): median device time per frame over `times`, incl. the image gather."""
    from .renderer_fused import FusedRenderer
    from . import parallel
    model = model or build_scene(device)
    model.eval()
    # a random-init field is transparent; scale sigma so rays terminate like a trained scene (sigma ~ 50 inside the figure)
    ro, rd = frame_rays(device)
    N = ro.shape[0]
    n_local = parallel.shard_tiles(N, world_size, rank).shape[0]
    fr = FusedRenderer(model, max_rays=n_local)
    out = {}
    ms_all = []
    for t in times:
        fr.render_sharded(ro, rd, float(t), rank, world_size, T_thresh=T_thresh)  # warm-up
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = fr.render_sharded(ro, rd, float(t), rank, world_size, T_thresh=T_thresh)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms_all.append(ts[len(ts) // 2])
        out["t=%.2f" % t] = {"ms": round(ts[len(ts) // 2], 3), "iterations": fr.iterations, "field_samples": fr.samples,
                             "coverage": round(float(res["weights_sum"].mean()), 4)}
    out["frame_ms_median"] = round(sorted(ms_all)[len(ms_all) // 2], 3)
    out["rays"] = N
    return out


def field_throughput(device, log2_M=20, reps=10, tf_peak=1678.0):
    """Fused field kernels at M = 2^log2_M samples: deform MLP / grid / heads forward (inference) TFLOP/s."""
    from . import field as F
    model = build_scene(device)
    cfg = model._field_cfg
    M = 1 << log2_M
    ws = F.FieldWorkspace(cfg, M, device, training=False)
    hw = F.HalfWeights(cfg, device)
    hw.refresh([w.detach() for w in model.mlp_weights()])
    table16 = model.encoder.embeddings.detach().half()
    xyz = torch.rand(M, 3, device=device) * 1.6 - 0.8
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, device=device), dim=-1)
    td = torch.tensor([0.4], device=device)
    st = _lib.stream

    def deform():
        F.deform_forward(cfg, hw, xyz, td, M, None, 1, ws.deform, ws.x01, None, None)

    def deform_mma_sync():
        _lib.call("seald_field_deform_forward", ptr(xyz), ptr(td), hw.p_deform, cfg.n_deform, M, None, cfg.bound, 1, ptr(ws.deform), ptr(ws.x01),
                  None, None, st())

    def heads():
        _lib.call("seald_field_heads_forward", ptr(ws.feat), ptr(dirs), hw.p_sigma, cfg.n_sigma, hw.p_color, cfg.n_color, M, None, 1.0,
                  ptr(ws.sigma), ptr(ws.rgb), None, None, None, None, st())

    F.field_forward(cfg, hw, ws, xyz, dirs, td, table16, model.encoder.offsets, None, 1)
    out = {"M": M}
    mac_deform = 76 * 128 + (cfg.n_deform - 2) * 128 * 128 + 128 * 3
    mac_heads = 32 * 64 + 64 * 16 + 31 * 64 + (cfg.n_color - 2) * 64 * 64 + 64 * 3
    for name, fn, fl in (("deform_fwd_tcgen05", deform, 2.0 * mac_deform * M), ("deform_fwd_mma_sync", deform_mma_sync, 2.0 * mac_deform * M),
                         ("heads_fwd", heads, 2.0 * mac_heads * M)):
        ms = _time(fn, reps=reps)
        tf = fl / (ms * 1e-3) / 1e12
        out[name] = {"ms": round(ms, 4), "TFLOP/s": round(tf, 1), "frac_of_tensor_peak": round(tf / tf_peak, 4)}
    return out


def main():
    dev = torch.device("cuda:0")
    _lib.load()
    peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    which = sys.argv[1:] or ["grid", "march"]
    if "grid" in which:
        for D in (3, 4):
            for gt in ("hash", "tiled"):
                print(json.dumps(grid_encoder(dev, 22, D, gt, torch.float16, hbm_gbs=hbm)), flush=True)
        print(json.dumps(grid_encoder(dev, 22, 3, "hash", torch.float16, coherent=True, hbm_gbs=hbm)), flush=True)
        print(json.dumps(grid_encoder(dev, 22, 3, "hash", torch.float32, hbm_gbs=hbm)), flush=True)
    if "sweep" in which:
        for lb in range(16, 25):
            r = grid_encoder(dev, lb, 3, "hash", torch.float16, hbm_gbs=hbm)
            print(json.dumps({"B": r["B"], "fwd": r["kernels"]["fwd"], "bwd": r["kernels"]["bwd_table_f32"]}), flush=True)
    if "march" in which:
        print(json.dumps(march_composite(dev, hbm_gbs=hbm)), flush=True)
    if "ncu" in which:
        # one short pass over the big-batch kernels (for an `ncu --set full` capture)
        print(json.dumps(grid_encoder(dev, 22, 3, "hash", torch.float16, reps=1, hbm_gbs=hbm)["kernels"]["fwd"]), flush=True)
        print(json.dumps(field_throughput(dev, reps=1)), flush=True)
        print(json.dumps(march_composite(dev, reps=1, hbm_gbs=hbm)), flush=True)
    if "field" in which:
        print(json.dumps(field_throughput(dev)), flush=True)
    if "frame" in which:
        print(json.dumps(frame_render(dev)), flush=True)


if __name__ == "__main__":
    main()
