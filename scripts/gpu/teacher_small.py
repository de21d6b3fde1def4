"""SealD teacher render of a 4096-ray batch: the exact round loop (packed rounds) vs the one-pass render (measurement script)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from seald_nerf_b200.renderer_fused import FusedRenderer
dev = torch.device("cuda:0")
teacher = bench.build_scene(dev, seed=0, seald=True); teacher.eval()
ro, rd, ts, gt = bench.make_batches(4, dev, 0)
for kind, mapper in bench.seald_mappers().items():
    teacher.init_mapper(mapper=mapper)
    for min_s in (1 << 16, 1 << 17, 1 << 18):
        fr = FusedRenderer(teacher, max_rays=4096, min_samples=min_s)
        def timeit(fn, K=30):
            for i in range(5): fn(i % 4)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K): fn(i % 4)
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / K
        a = timeit(lambda b: fr.render(ro[b], rd[b], ts[b], T_thresh=1e-4))
        it, sm = fr.iterations, fr.samples
        b_ = timeit(lambda b: fr.render_one_pass(ro[b], rd[b], ts[b], T_thresh=1e-4))
        o1 = fr.render(ro[0], rd[0], ts[0], T_thresh=1e-4)["image"].clone()
        o2 = fr.render_one_pass(ro[0], rd[0], ts[0], T_thresh=1e-4)["image"]
        print(kind, "min_samples", min_s, "round loop ms", round(a, 4), "rounds", it, "samples", sm, "| one pass ms", round(b_, 4), "samples", fr.samples,
              "| max diff", float((o1 - o2).abs().max()), flush=True)
    break
