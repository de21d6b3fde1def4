"""One rank's share of the 800x800 frame (interleaved 256-ray tiles of a W-way split) rendered on ONE GPU: what the round loop costs
as the per-rank ray count shrinks (measurement script; W from argv, default 1 2 4 8)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from seald_nerf_b200 import microbench, parallel
from seald_nerf_b200.renderer_fused import FusedRenderer
dev = torch.device("cuda:0")
model = microbench.build_scene(dev); model.eval()
ro, rd = microbench.frame_rays(dev)
N = ro.shape[0]
for W in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    idx = parallel.shard_tiles(N, W, 0).to(dev)
    o, d = ro.index_select(0, idx).contiguous(), rd.index_select(0, idx).contiguous()
    fr = FusedRenderer(model, max_rays=o.shape[0])
    for _ in range(2):
        fr.render(o, d, 0.5, T_thresh=1e-2)
    ts = []
    for _ in range(7):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fr.render(o, d, 0.5, T_thresh=1e-2); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("W", W, "rays", o.shape[0], "ms", round(ts[3], 3), "min", round(ts[0], 3), "rounds", fr.iterations, "samples", fr.samples, "launches", fr.launches,
          "cap", fr.cap, "slots", fr.slots, flush=True)
