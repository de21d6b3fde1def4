import sys, torch
sys.path.insert(0, "/root/repo")
from seald_nerf_b200 import microbench as mb
dev = torch.device("cuda:0")
from seald_nerf_b200.renderer_fused import FusedRenderer
model = mb.build_scene(dev); model.eval()
ro, rd = mb.frame_rays(dev)
fr = FusedRenderer(model, max_rays=ro.shape[0], use_graph=False)
fr.render(ro, rd, 0.5)
torch.cuda.synchronize()
print("MARK")
fr.render(ro, rd, 0.5)
torch.cuda.synchronize()
print(fr.iterations, fr.samples, fr.launches)
