"""Timing of the tcgen05 weight-gradient kernel at training-batch size, job subsets apart (measurement script)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from seald_nerf_b200 import microbench, _lib, field as F
dev = torch.device("cuda:0")
model = microbench.build_scene(dev)
cfg = model._field_cfg
M = 31715
Mp = 42368
ws = F.FieldWorkspace(cfg, Mp, dev, training=True)
for name in ("in_buf", "fwd_d", "bwd_d", "gout_d", "hs", "cin", "fwd_s", "fwd_c", "bwd_s", "bwd_c", "gout_s", "gout_c", "feat"):
    t = getattr(ws, name)
    t.copy_(torch.randn(t.shape, device=dev).to(t.dtype))
grads = [torch.zeros_like(w, dtype=torch.float32) for w in model.mlp_weights()]
m_dev = torch.tensor([M], dtype=torch.int32, device=dev)
jobs, n = F.wgrad_jobs(cfg, ws, grads, deform=True)
nd = cfg.n_deform
sub = {"all": (jobs, n), "deform": ((_lib.WgradJob * nd)(*[jobs[i] for i in range(nd)]), nd),
       "heads": ((_lib.WgradJob * (n - nd))(*[jobs[i] for i in range(nd, n)]), n - nd),
       "deform_mid": ((_lib.WgradJob * 6)(*[jobs[i] for i in range(1, 7)]), 6)}
out = {}
for name, (jb, k) in sub.items():
    fn = lambda: F.mlp_wgrad(jb, k, Mp, m_dev)
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    out[name] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
print("ROWW", os.environ.get("SEALD_WGRAD_ROWW", "default"), "us:", out, flush=True)
