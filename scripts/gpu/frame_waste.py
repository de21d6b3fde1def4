"""How many of the sample rows a frame's rounds hand to the field hold a sample (delta > 0)?  (measurement script)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from seald_nerf_b200 import microbench
from seald_nerf_b200.renderer_fused import FusedRenderer
dev = torch.device("cuda:0")
model = microbench.build_scene(dev); model.eval()
ro, rd = microbench.frame_rays(dev)
fr = FusedRenderer(model, max_rays=ro.shape[0], use_graph=False)
orig = fr._round
log = []
def wrapped(N, cur, first, opts, mapper, desc):
    st = fr.state.cpu()
    n = orig(N, cur, first, opts, mapper, desc)
    m = int(st[2])
    if m > 0:
        real = int((fr.deltas.view(-1, 2)[:m, 0] > 0).sum())
        log.append((int(st[0]), int(st[1]), m, real))
    return n
fr._round = wrapped
fr.render(ro, rd, 0.5, T_thresh=1e-2)
tot = sum(l[2] for l in log); real = sum(l[3] for l in log)
for l in log: print("alive %7d n_step %2d rows %8d real %8d (%.2f)" % (l + (l[3] / max(l[2], 1),)))
print("rows", tot, "real", real, "ratio", round(real / tot, 3))
