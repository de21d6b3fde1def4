import sys, time, torch
sys.path.insert(0, "/root/repo")
from seald_nerf_b200 import microbench as mb
from seald_nerf_b200.renderer_fused import FusedRenderer
dev = torch.device("cuda:0")
model = mb.build_scene(dev, seald=False); model.eval()
ro, rd = mb.frame_rays(dev)
g = torch.Generator(device="cpu").manual_seed(0)
idx = torch.randint(0, 640000, (4096,), generator=g).to(dev)
ro, rd = ro[idx].contiguous(), rd[idx].contiguous()
for graph in (True, False):
    for mns in (None,):
        fr = FusedRenderer(model, max_rays=4096, use_graph=graph)
        for _ in range(3):
            fr.render(ro, rd, 0.5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            fr.render(ro, rd, 0.5)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 20
        print("graph", graph, "max_n_step", mns, "ms", round(dt * 1e3, 3), "rounds", fr.iterations, "samples", fr.samples, "launches", fr.launches)
