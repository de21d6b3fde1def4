"""Density query (deform -> grid -> sigma) per-call time: one-launch tcgen05 kernel vs deformation kernel + grid/sigma kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from seald_nerf_b200 import field as F
dev = torch.device("cuda:0")
model = bench.build_scene(dev)
cfg = model._field_cfg
hw = model._half_weights(); hw.refresh([w.detach() for w in model.mlp_weights()]); hw.pack_sigma()
table16 = model.encoder.embeddings.detach().to(torch.float16)
H = 128
for M in (32768, 262144, 2 ** 21):
    # x-fastest lattice points like the occupancy sweep
    i = torch.arange(M, device=dev)
    xyz = torch.stack([(i % H), (i // H) % H, (i // (H * H)) % H], 1).float() / (H - 1) * 2 - 1
    xyz = (xyz + (torch.rand_like(xyz) * 2 - 1) / H).clamp(-1, 1).contiguous()
    ws = F.FieldWorkspace(cfg, M, dev, training=False)
    td = torch.tensor([0.4], device=dev)
    for impl in ("split", "umma"):
        F.DENSITY_IMPL = impl
        for _ in range(3):
            F.field_density(cfg, hw, ws, xyz, td, table16, model.encoder.offsets, sigma_packed=True, sigma_only=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            F.field_density(cfg, hw, ws, xyz, td, table16, model.encoder.offsets, sigma_packed=True, sigma_only=True)
        e1.record(); torch.cuda.synchronize()
        print("M", M, impl, "G", os.environ.get("SEALD_UMMA_G", "auto"), "ms", round(e0.elapsed_time(e1) / n, 4), flush=True)
