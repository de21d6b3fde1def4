"""Timing of the tcgen05 deformation kernels at training-batch size (measurement script)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from seald_nerf_b200 import microbench, _lib, field as F
dev = torch.device("cuda:0")
model = microbench.build_scene(dev)
cfg = model._field_cfg
hw = F.HalfWeights(cfg, dev)
hw.refresh([w.detach() for w in model.mlp_weights()])
td = torch.tensor([0.4], device=dev)
for M in (31711, 37888, 75776, 151552):
    Mp = (M + 127) // 128 * 128
    ws = F.FieldWorkspace(cfg, Mp, dev, training=True)
    xyz = torch.rand(Mp, 3, device=dev) * 1.6 - 0.8
    gx = torch.randn(Mp, 3, device=dev)
    m_dev = torch.tensor([M], dtype=torch.int32, device=dev)
    def fwd_save(): F.deform_forward(cfg, hw, xyz, td, Mp, m_dev, 1, ws.deform, ws.x01, ws.in_buf, ws.fwd_d)
    def fwd_nosave(): F.deform_forward(cfg, hw, xyz, td, Mp, m_dev, 1, ws.deform, ws.x01, None, None)
    def bwd(): F.deform_backward(cfg, hw, gx, td, Mp, m_dev, ws.fwd_d, ws.bwd_d, ws.gout_d)
    out = {}
    for name, fn in (("fwd_save", fwd_save), ("fwd_nosave", fwd_nosave), ("bwd", bwd)):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        out[name] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    print("G", os.environ.get("SEALD_UMMA_G", "auto"), "M", M, "us:", out, flush=True)
