"""Reference-vs-ours cases on one GPU: the REFERENCE's own host code (dnerf/network.py, dnerf/renderer.py, SealDNeRF/renderer.py,
their autograd wrappers) over its own CUDA extensions and cuBLAS `nn.Linear` under autocast — imported through
oracle/ref_runtime.py — beside this package's kernels, on the same weights, occupancy grid, rays and seeds.

Each `ref_*` function returns the reference's outputs, each `ours_*` the product's, both as dicts of CPU numpy arrays with the same
keys.  tests/test_gpu_ref_parity.py compares them live when the reference runtime travelled to the box (oracle/_ref/) and against
the committed fixtures tests/golden/ref_*.npz otherwise; tests/golden/make_ref_golden.py writes those fixtures from the `ref_*` side.
Pins SURVEY §8 rows a11, a12, a16, a17 (and f1) to something the reference computed.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

AUTOCAST = dict(device_type="cuda", dtype=torch.float16)
TABLE_SEED = 1


def np_(t):
    return t.detach().float().cpu().numpy()


def ours_model(dev, seed=0, seald=False):
    """Product model: random-init D-NeRF (reference initialisers, seed 0), analytic occupancy grid of the synthetic figure, and a
    hash table with structure (uniform +-0.5, CPU generator: the same numbers on every machine)."""
    import bench
    m = bench.build_scene(dev, seed, seald)
    g = torch.Generator().manual_seed(TABLE_SEED)
    with torch.no_grad():
        m.encoder.embeddings.copy_((torch.rand(m.encoder.embeddings.shape, generator=g) - 0.5).to(dev))
    return m


def ref_model(ours, seald=False):
    """The reference's network with the product model's parameters and buffers loaded through its own load_state_dict."""
    from oracle import ref_runtime as rr
    dev = ours.encoder.embeddings.device
    ref = rr.dnerf_network(seald=seald).to(dev)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref.mean_density, ref.iter_density, ref.mean_count, ref.local_step = ours.mean_density, ours.iter_density, ours.mean_count, ours.local_step
    return ref


def field_inputs(dev, n=4096, seed=3):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, 3, generator=g) * 1.8 - 0.9).to(dev)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
    return x, d


# ---- a12: NeRFNetwork.forward / .density ---------------------------------------------------------------------------------------
def _field(model, dev, ts=(0.37, 0.0)):
    x, d = field_inputs(dev)
    out = {}
    model.eval()
    for t in ts:
        tt = torch.tensor([[t]], dtype=torch.float32, device=dev)
        with torch.no_grad(), torch.autocast(**AUTOCAST):
            sigma, rgb, deform = model(x, d, tt)
            dens = model.density(x, tt)
        tag = "t%03d" % int(t * 100)
        out[tag + "_sigma"], out[tag + "_rgb"], out[tag + "_deform"] = np_(sigma), np_(rgb), np_(deform)
        out[tag + "_density_sigma"] = np_(dens["sigma"])
    return out


ref_field = _field
ours_field = _field


# ---- a17 training branch + a3/a7/a12 backward: image and every parameter gradient of one step -----------------------------------
def train_batch(dev, n=2048, seed=5):
    from helpers import camera_rays
    from seald_nerf_b200 import synthetic as syn
    ro, rd = camera_rays(n, seed=seed, center_crop=260)
    ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
    t = 20.5 / 64
    rgb, alpha = syn.render_gt(ro, rd, t, n_samples=192)
    return ro, rd, t, (rgb + (1 - alpha).unsqueeze(-1)).contiguous()


LOSS_SCALE = 8192.0  # (the reference trains with GradScaler's 65536 start; small scales push its fp16 gradients into subnormals)
TABLE_ROWS_KEPT = 65536  # dense levels 0-2 entirely + part of level 3; the other levels through per-level norms


def _grad_summary(gtable, offsets):
    """Fixture-sized view of a [rows, 2] table gradient: the first rows exactly + per-level sum / L2 norm / abs max."""
    g = gtable.detach().double()
    off = [int(o) for o in offsets]
    lv = []
    for l in range(len(off) - 1):
        s = g[off[l]:off[l + 1]]
        lv.append([float(s.sum()), float(s.pow(2).sum().sqrt()), float(s.abs().max())])
    return np_(gtable[:TABLE_ROWS_KEPT]), np.asarray(lv, np.float64)


def ref_train(ref, dev):
    ro, rd, t, gt = train_batch(dev)
    ref.train()
    ref.zero_grad(set_to_none=True)
    tt = torch.tensor([[t]], dtype=torch.float32, device=dev)
    with torch.autocast(**AUTOCAST):
        out = ref.render(ro[None], rd[None], tt, staged=False, bg_color=1, perturb=False, force_all_rays=False, dt_gamma=0, max_steps=1024)
        loss = ((out["image"] - gt[None]) ** 2).mean(-1).mean()      # criterion(pred, gt).mean(-1) ... .mean()  (dnerf/utils.py:82,111)
    (loss * LOSS_SCALE).backward()                                     # scaler.scale(loss).backward()            (nerf/utils.py:884)
    res = {"image": np_(out["image"][0]), "loss": np.float64(loss.item()), "samples": np.int64(int(ref.step_counter[(ref.local_step - 1) % 16, 0]))}
    head, lv = _grad_summary(ref.encoder.embeddings.grad / LOSS_SCALE, ref.encoder.offsets)
    res["grad_table_head"], res["grad_table_levels"] = head, lv
    for name, p in ref.named_parameters():
        if name != "encoder.embeddings":
            res["grad_" + name] = np_(p.grad / LOSS_SCALE)
    return res


def ours_train(ours, dev):
    from seald_nerf_b200.trainer import FusedTrainer
    ro, rd, t, gt = train_batch(dev)
    ours.train()
    tr = FusedTrainer(ours, num_rays=ro.shape[0], max_samples=ro.shape[0] * 160, use_graph=False, perturb=False, init_loss_scale=LOSS_SCALE)
    tr.set_inputs(ro, rd, t, gt)
    tr.grads.zero_()
    tr._forward_backward()
    torch.cuda.synchronize()
    res = {"image": np_(tr.pred), "loss": np.float64(tr.loss.item()),
           "samples": np.int64(int(tr.counter[0]))}
    head, lv = _grad_summary(tr.grad_table / LOSS_SCALE, ours.encoder.offsets)
    res["grad_table_head"], res["grad_table_levels"] = head, lv
    names = [n for n, _ in ours.named_parameters() if n != "encoder.embeddings"]
    by_param = {id(w): g for w, g in zip(ours.mlp_weights(), tr.grad_views)}
    for name, p in ours.named_parameters():
        if name != "encoder.embeddings":
            res["grad_" + name] = np_(by_param[id(p)] / LOSS_SCALE)
    assert len(names) == len(by_param)
    return res


# ---- f4: one step of Seal's local pre-training (SealNeRF/trainer.py:396-462) on the D-NeRF student --------------------------------
PRETRAIN_LR = 0.07  # init_pretraining default (SealDNeRF/utils.py:386)


def pretrain_batch(dev, n=4096, seed=13):
    """Lattice-free stand-in for a slice of pretraining_data: points inside an edited box, unit view directions, teacher labels."""
    g = torch.Generator().manual_seed(seed)
    pts = (torch.rand(n, 3, generator=g) * 0.6 - 0.3).to(dev)
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
    gt_sigma = (torch.rand(n, generator=g) * 40.0).to(dev)
    gt_color = torch.rand(n, 3, generator=g).to(dev)
    return pts, dirs, gt_sigma, gt_color, 0.37


def _pretrain_result(loss, gtable, offsets, table_after):
    head, lv = _grad_summary(gtable, offsets)
    return {"loss": np.float64(loss), "grad_table_head": head, "grad_table_levels": lv, "table_after_head": np_(table_after[:TABLE_ROWS_KEPT])}


def ref_pretrain(ref, dev):
    """pretrain_step's arithmetic on the reference network: L1Loss(sigma) + L1Loss(colour) under autocast, scaled backward, MLPs
    frozen (freeze_mlp; the SealD student's deformation net is frozen too), torch.optim.Adam on the table at the pre-training lr."""
    pts, dirs, gs, gc, t = pretrain_batch(dev)
    ref.train()
    ref.zero_grad(set_to_none=True)
    saved = ref.encoder.embeddings.detach().clone()
    for n_, p in ref.named_parameters():
        p.requires_grad_(n_ == "encoder.embeddings")
    tt = torch.tensor([[t]], dtype=torch.float32, device=dev)
    l1 = torch.nn.L1Loss()
    with torch.autocast(**AUTOCAST):
        out = ref(pts, dirs, tt)
        loss = l1(out[1], gc) * 1 + l1(out[0], gs)
    (loss * LOSS_SCALE).backward()
    g = ref.encoder.embeddings.grad / LOSS_SCALE       # scaler.unscale_
    opt = torch.optim.Adam([ref.encoder.embeddings], lr=PRETRAIN_LR, betas=(0.9, 0.99), eps=1e-15)
    ref.encoder.embeddings.grad = g.to(ref.encoder.embeddings.dtype)
    opt.step()
    res = _pretrain_result(loss.item(), g, ref.encoder.offsets, ref.encoder.embeddings)
    with torch.no_grad():
        ref.encoder.embeddings.copy_(saved)
    for p in ref.parameters():
        p.requires_grad_(True)
    ref.zero_grad(set_to_none=True)
    return res


def ours_pretrain(ours, dev):
    from seald_nerf_b200.trainer import FusedTrainer
    pts, dirs, gs, gc, t = pretrain_batch(dev)
    ours.train()
    saved = ours.encoder.embeddings.detach().clone()
    tr = FusedTrainer(ours, num_rays=1024, max_samples=pts.shape[0], use_graph=False, perturb=False, init_loss_scale=LOSS_SCALE, train_deform=False)
    tr.grads.zero_()
    tr.pretrain_step(pts, dirs, gs, gc, t, lr=PRETRAIN_LR, optimize=False)
    torch.cuda.synchronize()
    g = (tr.grad_table / LOSS_SCALE).clone()
    loss = float(tr.loss.item())
    tr._pretrain_optimize(PRETRAIN_LR)
    torch.cuda.synchronize()
    res = _pretrain_result(loss, g, ours.encoder.offsets, tr.params[:tr.n_table].view_as(ours.encoder.embeddings))
    res["table16_is_half_of_master"] = np.bool_(bool((tr.table16.reshape(-1)[:tr.n_table].float() == tr.params[:tr.n_table].half().float()).all()))
    with torch.no_grad():
        ours.encoder.embeddings.copy_(saved)
    return res


# ---- a17 eval branch: a whole 800x800 frame through the round loop --------------------------------------------------------------
FRAME_STRIDE = 16  # fixture keeps every 16th ray


def frame_rays(dev):
    from seald_nerf_b200 import microbench
    return microbench.frame_rays(dev, frame=3)


def ref_frame(ref, dev, t=0.5, T_thresh=None):
    ro, rd = frame_rays(dev)
    ref.eval()
    tt = torch.tensor([[t]], dtype=torch.float32, device=dev)
    kw = {} if T_thresh is None else {"T_thresh": T_thresh}
    with torch.no_grad(), torch.autocast(**AUTOCAST):
        out = ref.render(ro[None], rd[None], tt, staged=False, bg_color=1, perturb=False, dt_gamma=0, max_steps=1024, **kw)
    return {"image": np_(out["image"][0]), "depth": np_(out["depth"][0])}


def ours_frame(ours, dev, t=0.5, T_thresh=None):
    from seald_nerf_b200.renderer_fused import FusedRenderer
    ro, rd = frame_rays(dev)
    ours.eval()
    fr = FusedRenderer(ours, max_rays=ro.shape[0])
    out = fr.render(ro, rd, float(t), bg_color=1, perturb=False, dt_gamma=0, max_steps=1024, T_thresh=T_thresh)
    return {"image": np_(out["image"]), "depth": np_(out["depth"])}


def ours_frame_dropin(ours, dev, t=0.5):
    """The same frame through the drop-in NeRFRenderer.run_cuda (reference-shaped host loop over our ops)."""
    ro, rd = frame_rays(dev)
    ours.eval()
    tt = torch.tensor([[t]], dtype=torch.float32, device=dev)
    with torch.no_grad(), torch.autocast(**AUTOCAST):
        out = ours.render(ro[None], rd[None], tt, staged=False, bg_color=1, perturb=False, dt_gamma=0, max_steps=1024)
    return {"image": np_(out["image"][0]), "depth": np_(out["depth"][0])}


# ---- a11 / f1: update_extra_state (full sweep, then a partial pass) and mark_untrained_grid under the same seed ------------------
OCC_FRAMES_KEPT = (0, 21, 63)


def _occ_summary(m):
    g = m.density_grid
    res = {"mean_density": np.float64(m.mean_density), "occupied_per_frame": np_(torch.stack([b.to(torch.int32).ne(0).sum() for b in m.density_bitfield])),
           "grid_mean_per_frame": np_(g.clamp(min=0).mean(dim=(1, 2))), "untrained_cells": np.int64(int((g < 0).sum()))}
    bits = torch.stack([m.density_bitfield[t] for t in OCC_FRAMES_KEPT])
    res["bitfield_frames"] = bits.cpu().numpy()
    pop = torch.zeros(m.density_bitfield.shape[0], dtype=torch.int64)
    for t in range(m.density_bitfield.shape[0]):
        b = m.density_bitfield[t].to(torch.int32)
        pop[t] = sum(int(((b >> k) & 1).sum()) for k in range(8))
    res["occupied_cells_per_frame"] = pop.numpy()
    return res


def occ_poses(dev, n=24):
    from seald_nerf_b200 import synthetic as syn
    poses = syn.orbit_poses(n, dev, seed=2)
    fx, fy, cx, cy = syn.intrinsics()
    # a narrow field of view so that part of the grid is seen by no camera (otherwise nothing is marked)
    return poses, (fx * 6, fy * 6, cx, cy)


def run_occupancy(model, dev, fused, seed=11):
    """mark_untrained_grid, one full sweep, one partial pass.  `fused`: product pipeline (FusedOccupancy) instead of the host loop."""
    out = {}
    poses, intr = occ_poses(dev)
    model.density_grid.zero_()
    model.iter_density, model.local_step, model.mean_density = 0, 0, 0
    model.mark_untrained_grid(poses, intr)
    out["untrained_mask_frame0"] = np.packbits((model.density_grid[0, 0] < 0).cpu().numpy())
    occ = None
    if fused:
        from seald_nerf_b200.occupancy_fused import FusedOccupancy
        occ = FusedOccupancy(model)
    for tag, it in (("full", 0), ("partial", 16)):
        model.iter_density = it
        torch.manual_seed(seed)
        with torch.no_grad(), torch.autocast(**AUTOCAST):
            if occ is not None:
                occ.update()
            else:
                model.update_extra_state()
        torch.cuda.synchronize()
        for k, v in _occ_summary(model).items():
            out[tag + "_" + k] = v
        out[tag + "_grid_frame21"] = np_(model.density_grid[21, 0]).astype(np.float16)
    return out


# ---- a16: FFMLP against the reference's ffmlp extension --------------------------------------------------------------------------
def ffmlp_inputs(dev, B=4096, in_dim=32, seed=9):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, in_dim, generator=g) * 2 - 1).to(dev)
    go = torch.randn(B, 16, generator=g).to(dev)
    return x, go


def run_ffmlp(cls, dev, in_dim=32, out_dim=16, hidden=64, layers=2):
    """forward (training + inference) and backward of FFMLP(in, out, hidden, layers) with seeded weights."""
    torch.manual_seed(42)
    mlp = cls(in_dim, out_dim, hidden, layers).to(dev)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        mlp.weights.copy_(((torch.rand(mlp.weights.shape, generator=g) * 2 - 1) * (3.0 / hidden) ** 0.5).to(dev))
    x, go = ffmlp_inputs(dev, in_dim=in_dim)
    x = x.requires_grad_(True)
    mlp.train()
    with torch.autocast(**AUTOCAST):
        y = mlp(x)
    y.backward(go[:, :y.shape[1]].to(y.dtype))
    mlp.eval()
    with torch.no_grad(), torch.autocast(**AUTOCAST):
        y_inf = mlp(x.detach())
    torch.cuda.synchronize()
    return {"y": np_(y), "y_inference": np_(y_inf), "grad_x": np_(x.grad), "grad_w": np_(mlp.weights.grad)}


# ---- a17 SealD teacher (bbox / brush mapper) -------------------------------------------------------------------------------------
def seal_mapper_dict(kind):
    from oracle import seal as S
    if kind == "bbox":
        return S.make_bbox_mapper(center=(0.0, 0.15, 0.0), half=(0.15, 0.15, 0.15), translate=(0.2, 0.0, 0.0), rot_deg=30.0, hsv=(0.1, 0.0, 0.0))
    if kind == "brush_dry":
        return S.make_brush_mapper(mode="dry", rgb=(1.0, 0.0, 0.0))
    return S.make_brush_mapper(mode="linear")


def teacher_batch(dev, n=4096, seed=7):
    from helpers import camera_rays
    ro, rd = camera_rays(n, seed=seed, center_crop=300)
    return torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev), 0.3


def ref_teacher(ref, dev, kind):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_seal_golden import to_reference_mapper
    import SealNeRF.seal_utils as su
    mp = to_reference_mapper(su, seal_mapper_dict(kind))
    ref.init_mapper(config_dict={}, mapper=mp)
    ro, rd, t = teacher_batch(dev)
    ref.eval()
    tt = torch.tensor([[t]], dtype=torch.float32, device=dev)
    with torch.no_grad(), torch.autocast(**AUTOCAST):
        out = ref.render(ro[None], rd[None], tt, staged=False, bg_color=1, perturb=False, force_all_rays=True, dt_gamma=0, max_steps=1024,
                         T_thresh=1e-4)
    return {"image": np_(out["image"][0]), "depth": np_(out["depth"][0])}  # (the eval branch returns no weights_sum, SealDNeRF/renderer.py:286-289)


def ours_teacher(ours, dev, kind, one_pass=False):
    from helpers import seal_mapper_from_dict
    from seald_nerf_b200.renderer_fused import FusedRenderer
    ours.init_mapper(mapper=seal_mapper_from_dict(seal_mapper_dict(kind)))
    ro, rd, t = teacher_batch(dev)
    ours.eval()
    fr = FusedRenderer(ours, max_rays=ro.shape[0])
    fn = fr.render_one_pass if one_pass else fr.render
    out = fn(ro, rd, float(t), bg_color=1, T_thresh=1e-4)
    return {"image": np_(out["image"]), "depth": np_(out["depth"])}
