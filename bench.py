#!/usr/bin/env python
"""bench.py — D-NeRF hashgrid training throughput (BASELINE.json configs[1]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): one optimisation step of the D-NeRF field (hashgrid 16 levels x 2 features, T = 2^19,
8x128 time-conditioned deformation net, 64-wide sigma / colour heads) on a 4096-ray batch per GPU of the synthetic
jumpingjacks-shaped scene (800x800 cameras, t in [0,1]), `-O` semantics: fp16 tensor-core MLPs, cuda_ray march through
the occupancy bitfield, perturbed samples, loss-scaled Adam.  Data parallel across GPUs ("scaling": "weak"): every rank draws its own
4096-ray batch; the table / MLP gradients are summed by the fused exchange kernels of csrc/dp_fused.cu over NVLink peer memory
(symmetric memory + NVSwitch multicast; the reduce-scatter -> shard Adam -> fp16 broadcast is part of the step graph).  NCCL is
the process-group plumbing and the fallback exchange (SEALD_DP_MODE=sharded|allreduce).

`value`  : rays/s with the step's inputs already resident in HBM (device-timed with CUDA events, max over ranks).
`e2e`    : the same metric through the public API with HOST inputs (pinned H2D of rays/targets every step and a D2H
           read of the loss inside the timed region).
`roofline`: the dominant kernel of the step, timed alone with CUDA events, against MEASURED_PEAKS.json.
`frame`   : full 800x800 render (BASELINE configs[2]) through FusedRenderer, rays tile-sharded over the N GPUs + image
           all-gather, median device ms per frame (max over ranks).
`hashgrid`: GridEncoder 2^22 points x 16 levels forward / backward, GB/s of algorithmic bytes (BASELINE configs[4]).
`seald`   : SealD teacher->student distillation step (teacher render with the fused bbox proxy mapping + student train
           step with the frozen deformation net, BASELINE configs[3]) in rays/s.
`ref_gpu` : the REFERENCE's own CUDA path on this GPU (its host code + its extensions recompiled for sm_100a + cuBLAS autocast,
           oracle/ref_bench.py in a subprocess): train step, frame, occupancy refresh and per-kernel times beside ours.
`roofline_rows`: every north-star kernel (encoder fwd/bwd, march, composite, MLPs) at benchmark and at microbenchmark size.
`cpu_baseline` / `--impl reference`: the reference's pure-PyTorch path (cuda_ray=False, torch frequency encoders,
           BASELINE.json configs[0]) as ported in oracle/render.py, timed on the host cores (forward+render only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "D-NeRF train rays/s and 800x800 frame ms at 1/2/4/8 B200; hashgrid GB/s"
UNIT = "rays/s"
N_RAYS = 4096


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=10.0):
        """nvidia-smi takes a moment to start: block until its first sample so that short timed regions are covered."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        return time.perf_counter()

    def stop(self, t_begin=None, t_end=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.03)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if (t_begin is not None and ts < t_begin) or (t_end is not None and ts > t_end + 0.03):
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window": "device-timed loop + end-to-end loop"}


def run_reference(args, rank):
    """Reference arm: the reference's CPU implementation of the path (oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return
    from oracle import render
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    # bounded sample: each "step" renders ONE 4096-ray batch (forward+render, 128 steps/ray) ~1 s on 8 cores
    rps, cores, dt = render.time_cpu_render(n_rays=N_RAYS, num_steps=128, steps=min(steps, 10), warmup=warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rps, "unit": UNIT, "n_gpus": args.gpus, "steps": min(steps, 10), "warmup": warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "dnerf_cpu_pure_pytorch_forward_render_4096rays_128steps (BASELINE configs[0]; the reference has no CPU training path)"},
        "cpu_baseline": {"value": rps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d batches of 4096 rays x 128 samples, forward+render, fp32, torch %d threads" % (min(steps, 10), cores)},
        "e2e": {"value": rps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def build_scene(device, seed=0, seald=False):
    """Random-init D-NeRF (hashgrid) + analytic occupancy grid of the synthetic figure."""
    from seald_nerf_b200 import microbench
    model = microbench.build_scene(device, seed, seald)
    model.train()
    return model


def make_batches(n, device, rank, seed=0):
    """n training batches: (rays_o, rays_d [n,4096,3], times [n], gt [n,4096,3]) drawn like dnerf/provider.py:304-352."""
    import torch
    from seald_nerf_b200 import synthetic as syn
    intr = syn.intrinsics()
    poses = syn.orbit_poses(200, device, seed=seed)
    g = torch.Generator(device="cpu").manual_seed(1000 * rank + seed)
    ro, rd, ts, gt = [], [], [], []
    for i in range(n):
        frame = int(torch.randint(0, 200, (1,), generator=g))
        inds = torch.randint(0, 800 * 800, (N_RAYS,), generator=g).to(device)
        o, d = syn.get_rays(poses[frame], intr, 800, 800, inds)
        t = frame / 199.0
        rgb, alpha = syn.render_gt(o, d, t, n_samples=192)
        ro.append(o); rd.append(d); ts.append(t)
        gt.append(rgb + (1 - alpha).unsqueeze(-1))  # white background
    return torch.stack(ro), torch.stack(rd), ts, torch.stack(gt)


def seald_mappers():
    """The three edits of BASELINE configs[3], built from GUI-style configs by the package's own mapper construction."""
    import numpy as np
    from seald_nerf_b200.SealNeRF.seal_utils import get_seal_mapper
    # bbox tool: a 0.3^3 box on the torso, translated by +0.2 x and rotated 30 degrees about y
    c, h = np.array([0.0, 0.15, 0.0]), 0.15
    corners = np.array([[i, j, k] for i in (-1, 1) for j in (-1, 1) for k in (-1, 1)], np.float64) * h + c
    a = np.deg2rad(30.0)
    T = np.eye(4)
    T[:3, :3] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
    T[:3, 3] = np.array([0.2, 0.0, 0.0]) + c - T[:3, :3] @ c
    out = {"bbox": get_seal_mapper("", {"type": "bbox", "raw": corners.tolist(), "transform": T.tolist(), "scale": [1, 1, 1], "boundType": "to",
                                        "hsv": [0.1, 0.0, 0.0]})}
    # brush tool: two line strokes on the figure's front (plane x = 0.12), raised by 0.02 along +x, affecting 0.6 pressures of depth
    strokes = []
    for k in range(2):
        cc = np.array([0.12, 0.25 - 0.35 * k, 0.02 + 0.05 * k])
        u = np.linspace(-1, 1, 12)
        pts = [[0, p * 0.12, q * 0.05] for p in u for q in (-1, 1)] + [[0, q * 0.12, p * 0.05] for p in u for q in (-1, 1)] + \
              [[0, p * 0.06, q * 0.02] for p in (-1, 0, 1) for q in (-1, 1)]
        strokes.append((cc + np.array(pts)).tolist())
    for mode, extra in (("dry", {"rgb": [1.0, 0.0, 0.0]}), ("linear", {})):
        out["brush_" + mode] = get_seal_mapper("", dict({"type": "brush", "raw": strokes, "normal": [1, 0, 0], "brushType": "line", "brushDepth": 0.6,
                                                         "brushPressure": 0.02, "attenuationDistance": 0.02, "attenuationMode": mode}, **extra))
    return out


def seald_step_bench(device, rank, world, rays_o, rays_d, times, m_need, K=20):
    """SealD proxy-distillation step (SealDNeRF/utils.py:40-43,579-657 + the student step): the frozen teacher renders the
    batch through its Seal mapper (eval branch of SealNeRFTeacherRenderer.run_cuda), the student trains on that image.  One
    result per mapper of BASELINE configs[3]: bbox (headline), brush 'dry', brush 'linear'."""
    import torch
    from seald_nerf_b200.trainer import FusedTrainer
    from seald_nerf_b200.renderer_fused import FusedRenderer
    teacher = build_scene(device, seed=0, seald=True)
    teacher.eval()
    student = build_scene(device, seed=1, seald=True)
    trainer = FusedTrainer(student, num_rays=N_RAYS, max_samples=m_need, lr=1e-2, lr_net=1e-3, train_deform=False, world_size=world)
    n = rays_o.shape[0]
    res = {}
    for kind, mapper in seald_mappers().items():
        teacher.init_mapper(mapper=mapper)
        fr = FusedRenderer(teacher, max_rays=N_RAYS)

        def step(b):
            out = fr.render_one_pass(rays_o[b], rays_d[b], times[b], T_thresh=1e-4)
            trainer.train_step(rays_o[b], rays_d[b], times[b], torch.nan_to_num(out["image"]))

        for i in range(5):
            step(i % n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            step(i % n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        e0.record()
        for i in range(K):
            fr.render_one_pass(rays_o[i % n], rays_d[i % n], times[i % n], T_thresh=1e-4)
        e1.record()
        torch.cuda.synchronize()
        mask_count = int(fr.mask[:max(fr.samples, 1)].sum())  # mapped samples of the last teacher render
        res[kind] = {"ms_per_step": ms, "teacher_ms": e0.elapsed_time(e1) / K, "mapped_samples": mask_count, "teacher_iterations": fr.iterations}
        del fr
    trainer.flush()  # (data parallel: also a rank barrier — no rank frees its symmetric buffers while a peer's kernels may still read them)
    out = dict(res["bbox"])
    out["brush"] = {k[6:]: v for k, v in res.items() if k.startswith("brush_")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the frame / seald / hashgrid workloads (train step only)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-on-this-GPU subprocess (oracle/ref_bench.py)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from seald_nerf_b200.trainer import FusedTrainer
    from seald_nerf_b200 import _lib
    _lib.load()  # fails loudly without the CUDA library
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    K, W = args.steps, max(args.warmup, 3)

    model = build_scene(device)
    n_batches = min(K + W, 64)  # batches are cycled; 64 distinct ones keep set-up short
    rays_o, rays_d, times, gts = make_batches(n_batches, device, rank)
    # size the sample buffers like the reference's mean_count estimate (+25%)
    probe = FusedTrainer(model, num_rays=N_RAYS, max_samples=N_RAYS * 8, use_graph=False)
    m_need = max(probe.calibrate_max_samples(rays_o[i], rays_d[i], times[i]) for i in range(min(8, n_batches)))
    del probe
    if world > 1:
        mt = torch.tensor([m_need], device=device)
        dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        m_need = int(mt.item())
    trainer = FusedTrainer(model, num_rays=N_RAYS, max_samples=m_need, lr=1e-2, lr_net=1e-3, use_graph=not args.no_graph, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop --------------------------------------------------------------------------------
    for i in range(W):
        b = i % n_batches
        trainer.train_step(rays_o[b], rays_d[b], times[b], gts[b])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    t_clk0 = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        b = (W + i) % n_batches
        trainer.train_step(rays_o[b], rays_d[b], times[b], gts[b])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    loss_t = trainer.loss.clone()  # each rank holds its share of the global-batch mean (its squared errors / GLOBAL element count)
    if world > 1:
        dist.all_reduce(loss_t)
    loss_end = float(loss_t.item())
    launches = trainer.launches_per_step * K

    # ---- end-to-end loop (host inputs, loss read back every step) ------------------------------------------------
    # (a) synchronous: stage -> H2D -> step -> D2H loss -> host wait, every step; (b) software-pipelined (the headline `e2e`): the
    # next batch's H2D runs on a copy stream while the current step computes and the host reads each step's loss one step late —
    # the same bytes cross PCIe in both directions for every step, inside the timed region (FusedTrainer.train_step_host_pipelined).
    h_o, h_d, h_g = rays_o.cpu(), rays_d.cpu(), gts.cpu()
    for i in range(3):
        trainer.train_step_host(h_o[i % n_batches], h_d[i % n_batches], times[i % n_batches], h_g[i % n_batches])
    barrier()
    Ke = max(10, K // 2)
    t0 = time.perf_counter()
    e0.record()
    for i in range(Ke):
        b = i % n_batches
        trainer.train_step_host(h_o[b], h_d[b], times[b], h_g[b])
    e1.record()
    barrier()
    ms_e2e_sync = e0.elapsed_time(e1)
    for i in range(3):
        trainer.train_step_host_pipelined(h_o[i % n_batches], h_d[i % n_batches], times[i % n_batches], h_g[i % n_batches])
    trainer.drain_host_pipeline()
    barrier()
    e0.record()
    e2e_losses = []
    for i in range(Ke):
        b = i % n_batches
        e2e_losses.append(trainer.train_step_host_pipelined(h_o[b], h_d[b], times[b], h_g[b]))
    e2e_losses.append(trainer.drain_host_pipeline())  # the last step's loss is read inside the timed region too
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    assert all(l is not None and l == l for l in e2e_losses[1:])
    # ---- end-to-end with the training set resident on the GPU (the reference's `preload`): the step graph draws the pixels, builds the
    # rays and gathers the targets itself; per step the host sends a 4-byte frame index and reads the loss back ----
    ms_e2e_ds = 0.0
    if not args.no_extras:
        from seald_nerf_b200 import synthetic as syn
        F_ds = 8
        poses_all = syn.orbit_poses(200, device, seed=0)
        frames = [(25 * k + 7 * rank) % 200 for k in range(F_ds)]
        intr = syn.intrinsics()
        imgs = []
        for f in frames:
            o_f, d_f = syn.get_rays(poses_all[f], intr, 800, 800)
            parts = []
            for c0 in range(0, 800 * 800, 160000):
                rgb, alpha = syn.render_gt(o_f[c0:c0 + 160000], d_f[c0:c0 + 160000], f / 199.0, n_samples=192)
                parts.append(rgb + (1 - alpha).unsqueeze(-1))
            imgs.append(torch.cat(parts, 0))
        trainer.attach_dataset(poses_all[frames], intr, 800, 800, torch.stack(imgs), torch.tensor([f / 199.0 for f in frames]))
        del imgs
        for i in range(5):
            trainer.train_step_frame(i % F_ds, host_loss=True)
        barrier()
        e0.record()
        for i in range(Ke):
            trainer.train_step_frame(i % F_ds, host_loss=True)
        e1.record()
        barrier()
        ms_e2e_ds = e0.elapsed_time(e1)
    trainer.flush()  # the table pass of the last step's optimiser is deferred into the next step: apply it before the model is rendered
    steps_run = trainer.global_step
    # ---- replicas after all those steps: every rank must hold the same fp16 table and the same MLP weights (the fused exchange writes
    # each shard's refreshed rows into every rank's table; a lost or torn broadcast would show here) -------------------------------------
    dp_consistent = None
    if world > 1:
        trainer.sync_params()
        torch.cuda.synchronize()
        sums = torch.stack([trainer.table16.view(torch.int16).to(torch.int64).sum(), (trainer.table16.view(torch.int16).to(torch.int64)
                            * (torch.arange(trainer.table16.numel(), device=device) % 65521).view_as(trainer.table16)).sum(),
                            trainer.hw.flat.view(torch.int16).to(torch.int64).sum()])
        allsums = [torch.zeros_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        dp_consistent = {"ok": bool(all(torch.equal(a, allsums[0]) for a in allsums)),
                         "checksums_rank0": [int(v) for v in allsums[0].tolist()],
                         "what": "position-weighted integer checksums of the fp16 hash table and of the fp16 MLP weights after all steps, equal on every rank"}
    clocks = sampler.stop(t_clk0, sampler.mark()) if rank == 0 else None

    # ---- full-frame render, ray tiles sharded over the ranks (configs[2]) ------------------------------------------------
    from seald_nerf_b200 import microbench
    frame = None
    if not args.no_extras:
        model.eval()
        frame = microbench.frame_render(device, model=model, times=(0.0, 0.5, 1.0), reps=3, rank=rank, world_size=world)
        model.train()
    ms_frame = frame["frame_ms_median"] if frame else 0.0

    # ---- SealD distillation step: teacher render (bbox proxy mapping fused into the march) + student step (configs[3]) ----
    seald = None
    if not args.no_extras:
        seald = seald_step_bench(device, rank, world, rays_o, rays_d, times, m_need, K=max(20, K // 4))
    ms_seald = seald["ms_per_step"] if seald else 0.0

    # ---- occupancy-grid refresh (update_extra_state, every 100 steps in main_dnerf.py): fused device pipeline, frames sharded ----
    occ = None
    if not args.no_extras:
        saved = (model.density_grid.clone(), model.density_bitfield.clone(), model.iter_density, model.mean_density)
        occ = {}
        for name, it in (("full_sweep_ms", 0), ("partial_ms", 16)):
            model.iter_density = it
            trainer.update_extra_state()  # warm (allocations)
            model.iter_density = it
            barrier()
            t0 = time.perf_counter()
            trainer.update_extra_state()
            barrier()
            occ[name] = (time.perf_counter() - t0) * 1e3
        model.density_grid.copy_(saved[0]); model.density_bitfield.copy_(saved[1])
        model.iter_density, model.mean_density = saved[2], saved[3]
        trainer.refresh_occupancy()

    if world > 1:
        tt = torch.tensor([ms, ms_e2e, ms_frame, ms_seald, ms_e2e_ds, ms_e2e_sync], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_frame, ms_seald, ms_e2e_ds, ms_e2e_sync = (float(tt[i]) for i in range(6))

    # ---- per-stage timing + roofline of the dominant kernel (rank 0) ------------------------------------------------
    line = None
    if rank == 0:
        hbm, tf_burst, tf_sust, which = _peaks()
        trainer.set_inputs(rays_o[0], rays_d[0], times[0], gts[0])
        st = trainer.stage_timings(reps=20)
        m_live = st.pop("live_samples")
        cfg = trainer.cfg
        mac_deform = 76 * 128 + (cfg.n_deform - 2) * 128 * 128 + 128 * 3
        mac_heads = 32 * 64 + 64 * 16 + 31 * 64 + (cfg.n_color - 2) * 64 * 64 + 64 * 3
        def bytes_moved_by_adam(n_elems):
            """Bytes the table pass really moves: it reads g, m, v of every element (12 B) and, only where the gradient or a moment is
            non-zero, reads p and writes p, m, v and the fp16 copy (+18 B); the gradient is cleared where it was non-zero (+4 B)."""
            nt = (n_elems // 4) * 4
            touched = int((trainer.exp_avg[:nt].view(-1, 4) != 0).any(1).sum().item()) * 4
            return 12.0 * n_elems + 18.0 * touched + 4.0 * min(touched, 16 * 8 * m_live * cfg.grid_dim), touched

        n_opt = trainer.shard_len if trainer.dp_mode in ("fused", "sharded") else trainer.n_table_pad
        adam_moved, adam_touched = bytes_moved_by_adam(n_opt)
        work = {  # ALGORITHMIC work per launch (SURVEY.md §8d per-unit figures x live samples)
            "deform_fwd": ("tensor", 2.0 * mac_deform * m_live),
            "deform_bwd": ("tensor", 2.0 * (mac_deform - 76 * 128) * m_live),
            "wgrad": ("tensor", 2.0 * (mac_deform + mac_heads) * m_live),
            "heads_fwd": ("tensor", 2.0 * mac_heads * m_live),
            "heads_bwd": ("tensor", 2.0 * mac_heads * m_live),
            "grid_fwd": ("hbm", 588.0 * m_live),
            "grid_heads_fwd": ("hbm", 588.0 * m_live),   # encoder + heads in one launch, reported on the encoder's bytes
            "grid_scatter": ("hbm", 1100.0 * m_live),      # §8(d): 12 + 64 + 2 * 8 * 16 * 2 * 2 (fp16 table); the fp32 table we keep moves 2124
            "grid_input_bwd": ("hbm", (588.0 + 12.0) * m_live),
            "grid_bwd_both": ("hbm", (1100.0 + 588.0 + 12.0) * m_live),   # scatter + input gradient in one launch
            "march": ("hbm", 48.0 * N_RAYS + 32.0 * m_live + 262144.0),
            "composite_fwd": ("hbm", 24.0 * m_live + 32.0 * N_RAYS),
            "composite_bwd": ("hbm", 40.0 * m_live + 48.0 * N_RAYS),
            "composite_loss_fused": ("hbm", 64.0 * m_live + 104.0 * N_RAYS),
            # Adam over the fp32 table: 34 B/param if every element were updated; what the kernel really moves is computed above
            "optimizer": ("hbm", min(34.0 * trainer.n_params, adam_moved + 34.0 * trainer.n_weights)),
            "optimizer_table": ("hbm", min(34.0 * n_opt, adam_moved)),
        }
        if "optimizer_table" in st:
            work.pop("optimizer")  # there it is only the MLP-weight update + fp16 repack + loss-scale kernels
        else:
            work.pop("optimizer_table")

        def row(name, bound, amount, ms, size):
            sec = ms * 1e-3
            if bound == "tensor":
                a, pk, un = amount / sec / 1e12, tf_burst, "TFLOP/s"
            else:
                a, pk, un = amount / sec / 1e9, hbm, "GB/s"
            return {"kernel": name, "size": size, "bound": bound, "ms": round(ms, 5), "algorithmic": amount, "achieved": round(a, 2), "peak": pk,
                    "unit": un, "frac": round(a / pk, 4)}

        rows = [row(k, work[k][0], work[k][1], st[k], "train step: 4096 rays / %d samples" % m_live) for k in st if k in work and st[k] > 0]
        dom = max((k for k in st if k in work), key=lambda k: st[k])
        bound, amount = work[dom]
        sec = st[dom] * 1e-3
        if bound == "tensor":
            achieved, peak, unit = amount / sec / 1e12, tf_burst, "TFLOP/s"
        else:
            achieved, peak, unit = amount / sec / 1e9, hbm, "GB/s"
        # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture of the same step (profiles/, per round)
        traffic, traffic_src = None, None
        for tp in ("r2_traffic.json", "r1c_traffic.json"):
            tp = os.path.join(ROOT, "profiles", tp)
            if os.path.exists(tp) and traffic is None:
                tj = json.load(open(tp))
                ent = tj["per_launch_dram_bytes"].get(dom)
                if ent and world == 1:
                    traffic, traffic_src = ent["read"] + ent["write"], "%s: %s" % (tj["source"], ent["kernel"])
        roofline = {"kernel": dom, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak, "traffic": traffic,
                    "traffic_unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)", "traffic_source": traffic_src,
                    "algorithmic_per_launch": amount,
                    "bytes_note": ("optimizer: bytes the kernel really moves (12 B/element read + 22 B where a gradient or moment is non-zero; %d of %d "
                                   "elements live), not 34 B x every element" % (adam_touched, n_opt)) if dom.startswith("optimizer") else None,
                    "peak_source": which + (" burst" if bound == "tensor" else ""), "ms": st[dom],
                    "stage_ms": {k: round(v, 4) for k, v in st.items()}, "live_samples": m_live}

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import render
            rps, cores, dt = render.time_cpu_render(n_rays=N_RAYS, num_steps=128, steps=10, warmup=2)
            cpu = {"value": rps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "10 batches of 4096 rays x 128 samples, forward+render only (BASELINE configs[0]), fp32, %d torch threads" % cores}
        total_rays = N_RAYS * world
        ws_mb = (trainer.n_params * 18 + trainer.M * 4200) / 1e6
        line = {
            "dp_consistent": dp_consistent,
            "metric": METRIC, "value": total_rays * K / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "dnerf_hashgrid_L16_T19_F2_deform8x128_train_step_4096rays_per_gpu (BASELINE configs[1])",
                       "rays_per_gpu": N_RAYS, "max_samples": trainer.M, "live_samples": m_live, "cuda_graph": not args.no_graph,
                       "parallelism": "dp%d" % world, "final_loss": loss_end,
                       "l2": "no explicit flush: per-step working set (~%d MB of activations + optimiser state) exceeds the 126 MB L2" % ws_mb},
            "e2e": {"value": total_rays * Ke / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": trainer.h2d_bytes_per_step,
                    "d2h_bytes_per_step": 4, "steps": Ke, "ms_per_step": ms_e2e / Ke,
                    "mode": "software-pipelined: H2D of batch k+1 on a copy stream under step k, loss of step k read by the host after step k+1 "
                            "is enqueued; every step's inputs and loss cross PCIe inside the timed region",
                    "synchronous": {"value": total_rays * Ke / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync / Ke,
                                    "mode": "stage -> H2D -> step -> D2H loss -> host wait, every step (round 1's e2e)"}},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        }
        if ms_e2e_ds > 0:
            line["e2e_resident_dataset"] = {"value": total_rays * Ke / (ms_e2e_ds * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_ds / Ke, "steps": Ke,
                                            "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 4,
                                            "note": "8 training frames (800x800 RGB fp32) preloaded on the GPU like the reference's provider with "
                                                    "`preload`; pixel sampling + get_rays + target gather run inside the step graph"}
        if cpu:
            line["cpu_baseline"] = cpu
        if frame:
            line["frame"] = {"ms": ms_frame, "unit": "ms per 800x800 frame (max over ranks, incl. image all-gather)", "rays": frame["rays"],
                             "per_time": {k: v for k, v in frame.items() if k.startswith("t=")}, "n_gpus": world,
                             "sharding": "interleaved 256-ray tiles", "T_thresh": 1e-2}
        if seald:
            line["seald"] = {"value": N_RAYS * world / (ms_seald * 1e-3), "unit": "rays/s", "ms_per_step": ms_seald,
                             "teacher_ms": seald["teacher_ms"], "mapped_samples": seald["mapped_samples"],
                             "teacher_iterations": seald["teacher_iterations"],
                             "workload": "teacher eval render of the 4096-ray batch in one pass (bbox mapper 0.3^3, +0.2x, 30deg about y, fused in "
                                         "the march; FusedRenderer.render_one_pass) + student train step (frozen deform net)",
                             "brush": {k: dict(v, value=N_RAYS * world / (v["ms_per_step"] * 1e-3), unit="rays/s (this rank's step time)")
                                       for k, v in seald.get("brush", {}).items()},
                             "brush_workload": "the same step with the brush tool (two line strokes, pressure 0.02, depth 0.6): 'dry' recolours, "
                                               "'linear' pushes samples along the normal with border attenuation"}
        if occ:
            step_ms = ms / K
            line["occupancy_update"] = dict(occ, unit="wall ms per update_extra_state (64 time frames x 128^3 cells, sharded over the ranks, incl. its host syncs)",
                                            field_queries={"full_sweep": 64 * 128 ** 3, "partial": 64 * 128 ** 3 // 2},
                                            amortised_rays_per_s_every_100_steps=total_rays / ((step_ms + occ["partial_ms"] / 100.0) * 1e-3))
        line["config"]["skipped_steps"] = int(steps_run - int(trainer.step_dev.item())) if steps_run is not None else None
        line["config"]["skipped_steps_note"] = ("optimiser steps GradScaler skipped (overflow while the loss scale settles from 65536) out of %d run "
                                                "since the trainer was built; a skipped step does not touch Adam rows" % steps_run)
        if not args.no_extras:
            g = microbench.grid_encoder(device, 22, 3, "hash", torch.float16, reps=10, hbm_gbs=hbm)
            g4 = microbench.grid_encoder(device, 22, 4, "hash", torch.float16, reps=5, hbm_gbs=hbm)
            mc = microbench.march_composite(device, reps=5, hbm_gbs=hbm)
            fl = microbench.field_throughput(device, log2_M=20, reps=5, tf_peak=tf_burst)
            big = "2^22 uniform points"
            k3, k4 = g["kernels"], g4["kernels"]
            rows += [
                row("grid_fwd", "hbm", 588.0 * g["B"], k3["fwd"]["ms"], big + ", 3-D"),
                row("grid_fwd", "hbm", 1104.0 * g4["B"], k4["fwd"]["ms"], big + ", 4-D"),
                row("grid_bwd_table (fp32 gradient table, the product path)", "hbm", 1100.0 * g["B"], k3["bwd_table_f32"]["ms"], big + ", 3-D, §8d bytes (1100 B/pt)"),
                row("grid_bwd_table (fp32 gradient table, the product path)", "hbm", 2124.0 * g["B"], k3["bwd_table_f32"]["ms"], big + ", 3-D, bytes of the fp32 table (2124 B/pt)"),
                row("grid_bwd_table (fp16 gradient table, the reference's dtype)", "hbm", 1100.0 * g["B"], k3["bwd_table_same_dtype"]["ms"], big + ", 3-D"),
                row("march_train", "hbm", mc["march_train"]["algorithmic_GB"] * 1e9, mc["march_train"]["ms"], "%d rays / %d samples" % (mc["rays"], mc["samples"])),
                row("composite_fwd", "hbm", mc["composite_fwd"]["algorithmic_GB"] * 1e9, mc["composite_fwd"]["ms"], "%d rays / %d samples" % (mc["rays"], mc["samples"])),
                row("composite_bwd", "hbm", mc["composite_bwd"]["algorithmic_GB"] * 1e9, mc["composite_bwd"]["ms"], "%d rays / %d samples" % (mc["rays"], mc["samples"])),
                row("packbits", "hbm", mc["packbits_64frames"]["algorithmic_GB"] * 1e9, mc["packbits_64frames"]["ms"], "64 frames x 128^3 cells"),
                row("deform_fwd (tcgen05)", "tensor", 2.0 * mac_deform * fl["M"], fl["deform_fwd_tcgen05"]["ms"], "2^20 samples"),
                row("heads_fwd", "tensor", 2.0 * mac_heads * fl["M"], fl["heads_fwd"]["ms"], "2^20 samples"),
            ]
            line["roofline_rows"] = rows
            line["hashgrid_4d"] = {"points": g4["B"], "fwd_ms": k4["fwd"]["ms"], "fwd_GBs": k4["fwd"]["GB/s"], "fwd_frac_of_hbm": k4["fwd"]["frac_of_hbm"],
                                   "bwd_ms": k4["bwd_table_f32"]["ms"], "bytes_per_point": {"fwd": 1104}}
            line["hashgrid"] = {"points": g["B"], "levels": 16, "table": "2^19 x 2 fp16", "fwd_GBs": g["kernels"]["fwd"]["GB/s"],
                                "fwd_frac_of_hbm": g["kernels"]["fwd"]["frac_of_hbm"], "fwd_ms": g["kernels"]["fwd"]["ms"],
                                "bwd_GBs": g["kernels"]["bwd_table_f32"]["GB/s"], "bwd_frac_of_hbm": g["kernels"]["bwd_table_f32"]["frac_of_hbm"],
                                "bwd_ms": g["kernels"]["bwd_table_f32"]["ms"], "bytes_per_point": {"fwd": 588, "bwd_f32_table": 2124},
                                "bwd_frac_of_hbm_on_1100B": round(1100.0 * g["B"] / (g["kernels"]["bwd_table_f32"]["ms"] * 1e-3) / 1e9 / hbm, 4)}
        else:
            line["roofline_rows"] = rows
        if world == 1 and not args.no_ref_gpu and not args.no_extras:
            # the reference's own CUDA path on this GPU, in its own process (it launches on the legacy default stream)
            try:
                pr = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py"), "--steps", str(min(max(K, 20), 100)),
                                     "--warmup", "20"], capture_output=True, text=True, timeout=600)
                rg = json.loads(pr.stdout.strip().splitlines()[-1]) if pr.returncode == 0 and pr.stdout.strip() else {
                    "unavailable": "oracle/ref_bench.py rc %d: %s" % (pr.returncode, pr.stderr[-300:])}
            except Exception as e:  # noqa: BLE001
                rg = {"unavailable": repr(e)[:300]}
            if "train" in rg:
                rg["ours_over_ref"] = {
                    "train_step": round(rg["train"]["ms_per_step"] / (ms / K), 2),
                    "train_step_e2e": round(rg["train"]["e2e_ms_per_step"] / (ms_e2e / Ke), 2),
                    "frame": round(rg["frame"]["frame_ms_median"] / ms_frame, 2) if ms_frame > 0 and "frame" in rg else None,
                    "occupancy_full": round(rg["occupancy_update"]["full_sweep_ms"] / occ["full_sweep_ms"], 2) if occ and "occupancy_update" in rg else None,
                    "occupancy_partial": round(rg["occupancy_update"]["partial_ms"] / occ["partial_ms"], 2) if occ and "occupancy_update" in rg else None,
                    "grid_fwd_2p22": round(rg["kernels"]["grid_3d_fwd"]["ms"] / line["hashgrid"]["fwd_ms"], 2) if "kernels" in rg else None,
                    "grid_bwd_2p22": round(rg["kernels"]["grid_3d_bwd_fp16_table"]["ms"] / line["hashgrid"]["bwd_ms"], 2) if "kernels" in rg else None,
                }
            line["ref_gpu"] = rg
        print(json.dumps(line), flush=True)
    if world > 1:
        # Teardown with a bounded wait (a communicator teardown that blocks must not turn a finished measurement into a hang).
        # Order matters: the barrier comes FIRST — rank 0 may still be timing stages whose kernels touch the peers' symmetric
        # memory, so no rank may free its buffers before every rank got here; then the CUDA graphs holding captured exchange
        # kernels are released, then the communicator.
        holder = [trainer]
        del trainer

        def _teardown():
            dist.barrier()
            holder.clear()
            torch.cuda.synchronize()
            dist.destroy_process_group()
        th = threading.Thread(target=_teardown, daemon=True)
        th.start()
        th.join(timeout=60)
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
