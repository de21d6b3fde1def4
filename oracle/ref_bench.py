"""TEST / MEASUREMENT INFRASTRUCTURE — the REFERENCE's CUDA path timed on this GPU ("ref-on-B200").

The reference's own host code (dnerf/network.py + dnerf/renderer.py `run_cuda`, its autograd wrappers, torch.optim.Adam +
GradScaler exactly as nerf/utils.py:879-886 / main_dnerf.py:129 drive them) over its own extensions recompiled for sm_100a
(oracle/_ref/) and cuBLAS under autocast — nothing of this package's kernels is on the timed path.  This package only supplies the
synthetic scene (analytic occupancy grid, rays, targets), which is data, so both arms see the same inputs.

    python oracle/ref_bench.py [--steps K] [--warmup W] [--what train,frame,kernels,occupancy]

prints ONE JSON line.  bench.py runs it in a subprocess (rank 0, N = 1) and embeds the result as `ref_gpu`.
CUDA events on the reference's stream (it launches on the legacy default stream = torch's current stream), warm-up first.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_RAYS = 4096


def _median_ms(fn, reps, warmup=3, flush=None):
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


def build_reference_scene(dev, seed=0):
    """Reference network, random init (its own initialisers), with the analytic occupancy grid of the synthetic figure packed by the
    reference's own packbits."""
    from oracle import ref_runtime as rr
    from seald_nerf_b200 import synthetic as syn
    rr.install()
    import raymarching
    torch.manual_seed(seed)
    net = rr.dnerf_network().to(dev)
    grid = syn.make_density_grid(net.time_size, net.grid_size, 1.0, dev)
    net.density_grid.copy_(grid)
    net.mean_density = float(grid.clamp(min=0).mean())
    thresh = min(net.mean_density, net.density_thresh)
    for t in range(net.time_size):
        raymarching.packbits(net.density_grid[t], thresh, net.density_bitfield[t])
    return net


def train_step_bench(net, dev, steps, warmup):
    """nerf/utils.py:879-886 with dnerf/utils.py:38-115's train_step (C == 3 images: bg_color = 1), Adam of main_dnerf.py:129."""
    import bench
    n_b = min(steps + warmup, 64)
    rays_o, rays_d, times, gts = bench.make_batches(n_b, dev, 0)
    net.train()
    opt = torch.optim.Adam(net.get_params(1e-2, 1e-3), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda", enabled=True)
    kw = dict(staged=False, bg_color=1, perturb=True, force_all_rays=False, dt_gamma=0, max_steps=1024)

    def step(b):
        opt.zero_grad()
        tt = torch.tensor([[times[b]]], dtype=torch.float32, device=dev)
        with torch.autocast("cuda", dtype=torch.float16):
            out = net.render(rays_o[b][None], rays_d[b][None], tt, **kw)
            loss = ((out["image"] - gts[b][None]) ** 2).mean(-1).mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss

    for i in range(warmup):
        step(i % n_b)
        if i == min(15, warmup - 1):
            # what update_extra_state does every 16 steps (dnerf/renderer.py:551-554): the march then sizes its buffers from mean_count
            # and stops synchronising (raymarching.py:196-221)
            net.mean_count = int(net.step_counter[:min(16, net.local_step), 0].sum().item() / max(1, min(16, net.local_step)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step((warmup + i) % n_b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # end to end: host batches in (pinned H2D), loss read back every step (what Trainer.train_one_epoch does with loss.item())
    h_o, h_d, h_g = rays_o.cpu().pin_memory(), rays_d.cpu().pin_memory(), gts.cpu().pin_memory()
    ke = max(10, steps // 2)

    def step_host(b):
        rays_o[b].copy_(h_o[b], non_blocking=True); rays_d[b].copy_(h_d[b], non_blocking=True); gts[b].copy_(h_g[b], non_blocking=True)
        return float(step(b).item())

    for i in range(3):
        step_host(i % n_b)
    torch.cuda.synchronize()
    e0.record()
    for i in range(ke):
        step_host(i % n_b)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / ke
    return {"ms_per_step": ms, "rays_per_s": N_RAYS / (ms * 1e-3), "e2e_ms_per_step": ms_e2e, "e2e_rays_per_s": N_RAYS / (ms_e2e * 1e-3),
            "steps": steps, "warmup": warmup, "mean_count": int(net.mean_count), "final_loss": float(loss.item()),
            "samples_last_step": int(net.step_counter[(net.local_step - 1) % 16, 0])}


def frame_bench(net, dev, times=(0.0, 0.5, 1.0), reps=3):
    from seald_nerf_b200 import microbench
    ro, rd = microbench.frame_rays(dev)
    net.eval()
    out = {}
    ms_all = []
    for T_thresh_tag, kw in (("T1e-2", {}),):
        for t in times:
            tt = torch.tensor([[t]], dtype=torch.float32, device=dev)

            def fn():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                    return net.render(ro[None], rd[None], tt, staged=False, bg_color=1, perturb=False, dt_gamma=0, max_steps=1024, **kw)
            ms = _median_ms(fn, reps, warmup=1)
            out["t=%.2f" % t] = round(ms, 3)
            ms_all.append(ms)
    out["frame_ms_median"] = round(sorted(ms_all)[len(ms_all) // 2], 3)
    out["rays"] = ro.shape[0]
    return out


def kernel_bench(dev, hbm_gbs):
    """The reference's extensions called directly (the pybind entry points its wrappers call), same inputs as seald_nerf_b200.microbench."""
    from oracle import ref_runtime as rr
    from seald_nerf_b200 import synthetic as syn, microbench
    ge, rm = rr.load_extension("gridencoder"), rr.load_extension("raymarching")
    out = {}
    # ---- grid encoder 2^22 points, 3-D and 4-D, L16 F2 T2^19 fp16 (configs[4]) -------------------------------------------------
    rr.install()
    from gridencoder.grid import GridEncoder
    for D in (3, 4):
        B, L, C = 1 << 22, 16, 2
        enc = GridEncoder(input_dim=D, num_levels=L, level_dim=C, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048).to(dev)
        g = torch.Generator(device=dev).manual_seed(0)
        x = torch.rand(B, D, device=dev, generator=g)
        table = enc.embeddings.detach().half().uniform_(-0.1, 0.1)
        outp = torch.empty(L, B, C, device=dev, dtype=torch.half)
        dy_dx = torch.empty(B, L * D * C, device=dev, dtype=torch.half)
        grad = torch.randn(L, B, C, device=dev).half()
        gtab = torch.zeros_like(table)
        gx = torch.zeros(B, D, device=dev, dtype=torch.half)
        S, H = float(np.log2(enc.per_level_scale)), 16
        a = (enc.gridtype_id, enc.align_corners, enc.interp_id) if hasattr(enc, "interp_id") else (enc.gridtype_id, enc.align_corners)

        def fwd():
            ge.grid_encode_forward(x, table, enc.offsets, outp, B, D, C, L, S, H, None, *a)

        def fwd_dydx():
            ge.grid_encode_forward(x, table, enc.offsets, outp, B, D, C, L, S, H, dy_dx, *a)

        def bwd():
            ge.grid_encode_backward(grad, x, table, enc.offsets, gtab, B, D, C, L, S, H, None, None, *a)

        def bwd_dx():
            ge.grid_encode_backward(grad, x, table, enc.offsets, gtab, B, D, C, L, S, H, dy_dx, gx, *a)

        nc = 1 << D
        b_fwd = B * (4 * D + nc * L * C * 2 + L * C * 2)
        b_bwd = B * (4 * D + L * C * 2 + 2 * nc * L * C * 2)
        for name, fn, nb in (("fwd", fwd, b_fwd), ("fwd_dy_dx", fwd_dydx, b_fwd + B * L * D * C * 2), ("bwd_fp16_table", bwd, b_bwd),
                             ("bwd_fp16_table+input_grad", bwd_dx, b_bwd + B * (L * D * C * 2 + 4 * D))):
            ms = _median_ms(fn, 10)
            out["grid_%dd_%s" % (D, name)] = {"ms": round(ms, 4), "GB/s": round(nb / ms / 1e6, 1), "frac_of_hbm": round(nb / ms / 1e6 / hbm_gbs, 4)}
        # wrapper-level forward (incl. the reference's permute + contiguous copy of [L,B,C] -> [B,L*C], grid.py:57)
        xr = (x * 2 - 1).requires_grad_(False)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            ms = _median_ms(lambda: enc(xr, bound=1), 10)
        out["grid_%dd_module_forward" % D] = {"ms": round(ms, 4)}
        del enc, x, table, outp, dy_dx, grad, gtab, gx
    # ---- march / composite on a whole frame (640 000 rays) ----------------------------------------------------------------------
    Hg = 128
    grid = syn.make_density_grid(64, Hg, 1.0, dev)
    bits = torch.empty(Hg ** 3 // 8, dtype=torch.uint8, device=dev)
    rm.packbits(grid[20].contiguous(), Hg ** 3 // 8, 10.0, bits)
    ro, rd = microbench.frame_rays(dev)
    N = ro.shape[0]
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=dev)
    nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
    rm.near_far_from_aabb(ro, rd, aabb, N, 0.2, nears, fars)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    noises = torch.zeros(N, device=dev)
    M0 = N * 16
    xyzs = torch.zeros(M0, 3, device=dev); dirs = torch.zeros(M0, 3, device=dev); deltas = torch.zeros(M0, 2, device=dev)
    rays = torch.zeros(N, 3, dtype=torch.int32, device=dev)
    rm.march_rays_train(ro, rd, bits, 1.0, 0.0, 1024, N, 1, Hg, M0, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
    m_live = int(counter[0])
    M = (m_live + 127) // 128 * 128
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def march():
        counter.zero_()
        # the wrapper zero-fills the three sample buffers before every launch (raymarching.py:205-207)
        xyzs[:M].zero_(); dirs[:M].zero_(); deltas[:M].zero_()
        rm.march_rays_train(ro, rd, bits, 1.0, 0.0, 1024, N, 1, Hg, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises)

    sig = torch.rand(M, device=dev) * 30
    rgb = torch.rand(M, 3, device=dev)
    ws = torch.empty(N, device=dev); depth = torch.empty(N, device=dev); image = torch.empty(N, 3, device=dev)
    gws = torch.rand(N, device=dev); gim = torch.rand(N, 3, device=dev)
    gs = torch.zeros(M, device=dev); gc = torch.zeros(M, 3, device=dev)

    def comp_fwd():
        rm.composite_rays_train_forward(sig, rgb, deltas, rays, M, N, 1e-4, ws, depth, image)

    def comp_bwd():
        gs.zero_(); gc.zero_()  # raymarching.py:283-284
        rm.composite_rays_train_backward(gws, gim, sig, rgb, deltas, rays, ws, image, M, N, 1e-4, gs, gc)

    gall = grid.reshape(64, -1).contiguous()
    ball = torch.empty(64, Hg ** 3 // 8, dtype=torch.uint8, device=dev)

    def pack_all():
        for t in range(64):  # dnerf/renderer.py:547-548: one launch per time frame
            rm.packbits(gall[t], Hg ** 3 // 8, 10.0, ball[t])

    out["frame_rays"], out["frame_samples"] = N, m_live
    for name, fn, nb in (("march_train", march, 48.0 * N + 32.0 * m_live + 262144.0), ("composite_fwd", comp_fwd, 24.0 * m_live + 32.0 * N),
                         ("composite_bwd", comp_bwd, 40.0 * m_live + 48.0 * N), ("packbits_64frames", pack_all, 4.125 * 64 * Hg ** 3)):
        ms = _median_ms(fn, 10, flush=flush)
        out[name] = {"ms": round(ms, 4), "GB/s": round(nb / ms / 1e6, 1), "frac_of_hbm": round(nb / ms / 1e6 / hbm_gbs, 4)}
    return out


def occupancy_bench(net, dev):
    import time
    out = {}
    saved = (net.density_grid.clone(), net.density_bitfield.clone(), net.mean_density)
    for name, it in (("full_sweep_ms", 0), ("partial_ms", 16)):
        net.iter_density = it
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.float16):  # nerf/utils.py:871-872
            net.update_extra_state()
        torch.cuda.synchronize()
        out[name] = (time.perf_counter() - t0) * 1e3
    net.density_grid.copy_(saved[0]); net.density_bitfield.copy_(saved[1]); net.mean_density = saved[2]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--what", default="train,frame,kernels,occupancy")
    args = ap.parse_args()
    what = args.what.split(",")
    from oracle import ref_runtime as rr
    if not rr.available():
        print(json.dumps({"unavailable": "oracle/_ref (reference extensions + byte-compiled host modules) not present"}))
        return
    dev = torch.device("cuda:0")
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks)).get("hbm_gbs", 6650.0) if os.path.exists(peaks) else 6650.0
    res = {"what": "reference CUDA path on this GPU: reference host code + its extensions recompiled for sm_100a + cuBLAS autocast",
           "gpu": torch.cuda.get_device_name(0)}
    net = build_reference_scene(dev)
    if "train" in what:
        res["train"] = train_step_bench(net, dev, args.steps, max(args.warmup, 17))
    if "frame" in what:
        res["frame"] = frame_bench(net, dev)
    if "occupancy" in what:
        res["occupancy_update"] = occupancy_bench(net, dev)
    if "kernels" in what:
        del net
        torch.cuda.empty_cache()
        res["kernels"] = kernel_bench(dev, hbm)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
