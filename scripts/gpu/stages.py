"""Per-stage timing of the fused training step at the benchmark size (measurement script): python scripts/gpu/stages.py [stage ...]"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from seald_nerf_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
model = bench.build_scene(dev)
ro, rd, ts, gt = bench.make_batches(4, dev, 0)
tr = FusedTrainer(model, num_rays=4096, max_samples=42368, lr=1e-2, lr_net=1e-3)
for i in range(8):
    tr.train_step(ro[i % 4], rd[i % 4], ts[i % 4], gt[i % 4])
torch.cuda.synchronize()
res = {}
for rep in range(3):
    st = tr.stage_timings()
    for k, v in st.items():
        res.setdefault(k, []).append(v)
want = sys.argv[1:]
out = {k: round(min(v), 4) for k, v in res.items() if not want or k in want}
print(json.dumps(out), flush=True)
