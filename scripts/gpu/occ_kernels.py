"""Warm, event-timed kernels of ONE occupancy frame (2^21 lattice points) for different batch orders (measurement script)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from seald_nerf_b200 import _lib, field as F
from seald_nerf_b200._lib import ptr
dev = torch.device("cuda:0")
model = bench.build_scene(dev)
cfg = model._field_cfg
train_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if train_steps:
    from seald_nerf_b200.trainer import FusedTrainer
    tr = FusedTrainer(model, num_rays=4096, max_samples=42368, lr=1e-2, lr_net=1e-3)
    ro, rd, ts, gt = bench.make_batches(4, dev, 0)
    for i in range(train_steps):
        tr.train_step(ro[i % 4], rd[i % 4], ts[i % 4], gt[i % 4])
    tr.flush(); torch.cuda.synchronize()
    print("trained", train_steps, "steps", flush=True)
hw = model._half_weights(); hw.refresh([w.detach() for w in model.mlp_weights()])
table16 = model.encoder.embeddings.detach().half()
H = 128; n = H ** 3
ws = F.FieldWorkspace(cfg, n, dev, training=False)
rnd = torch.rand(n, 3, device=dev)
xyz = torch.empty(n, 3, device=dev); idx = torch.empty(n, dtype=torch.int32, device=dev)
_lib.call("seald_occ_cell_points", None, ptr(rnd), n, H, float(1 - 1 / H), float(1 / H), ptr(xyz), ptr(idx), _lib.stream())
td = torch.tensor([0.37], device=dev)
orders = {"x_fastest": xyz.clone(), "z_fastest": xyz.view(H, H, H, 3).permute(2, 1, 0, 3).reshape(-1, 3).contiguous(),
          "morton": xyz[torch.argsort(idx.long())].contiguous(), "random": xyz[torch.randperm(n, device=dev)].contiguous()}
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 4)
off = model.encoder.offsets
for name, p in orders.items():
    d = lambda: F.deform_forward(cfg, hw, p, td, n, None, 2, ws.deform, ws.x01, None, None)
    g = lambda: _lib.call("seald_grid_encode_forward", ptr(ws.x01), ptr(table16), ptr(off), ptr(ws.feat), None, n, 3, cfg.grid_dim, cfg.grid_levels,
                          cfg.grid_S, cfg.grid_base, cfg.gridtype, int(cfg.align_corners), cfg.interp, _lib.F16, None, _lib.stream())
    s = lambda: _lib.call("seald_field_sigma_forward", ptr(ws.feat), hw.p_sigma, cfg.n_sigma, n, 1.0, ptr(ws.sigma), None, _lib.stream())
    print(name, "ms: deform", timeit(d), "grid", timeit(g), "sigma", timeit(s), "| oob frac", float(((ws.x01 < 0) | (ws.x01 > 1)).any(-1).float().mean()),
          "mean |dx|", float(ws.deform.abs().mean()), flush=True)
