"""Per-entry-point device time of one 800x800 frame (eager rounds, every C-ABI call bracketed by events; measurement script)."""
import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from seald_nerf_b200 import microbench, _lib
from seald_nerf_b200.renderer_fused import FusedRenderer
dev = torch.device("cuda:0")
model = microbench.build_scene(dev); model.eval()
ro, rd = microbench.frame_rays(dev)
fr = FusedRenderer(model, max_rays=ro.shape[0], use_graph=False)
fr.render(ro, rd, 0.5, T_thresh=1e-2)
acc = collections.OrderedDict()
orig = _lib.call
def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(name, *a); e1.record(); torch.cuda.synchronize()
    x = acc.setdefault(name, [0, 0.0]); x[0] += 1; x[1] += e0.elapsed_time(e1)
    if "march" in name:
        st = fr.state.cpu().tolist()
        print("  ", name, "alive", st[0], "n_step", st[1], "rows", st[6] if fr.pack else st[2], "ms", round(e0.elapsed_time(e1), 4))
    return r
_lib.call = timed
fr.render(ro, rd, 0.5, T_thresh=1e-2)
_lib.call = orig
print("pack", fr.pack, "rounds", fr.iterations, "samples", fr.samples)
for k, (n, ms) in acc.items():
    print("%-40s %3d %8.3f ms" % (k, n, ms))
print("sum", round(sum(v[1] for v in acc.values()), 3))
