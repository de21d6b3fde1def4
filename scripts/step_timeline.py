"""In-graph timeline of one training step on one GPU (FusedTrainer.step_timeline) next to the isolated stage timings."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from seald_nerf_b200.trainer import FusedTrainer
dev = torch.device("cuda", 0)
model = bench.build_scene(dev)
ro, rd, ts, gt = bench.make_batches(4, dev, 0)
tr = FusedTrainer(model, num_rays=4096, max_samples=42368, lr=1e-2, lr_net=1e-3)
for i in range(20):
    tr.train_step(ro[i % 4], rd[i % 4], ts[i % 4], gt[i % 4])
tr.set_inputs(ro[0], rd[0], ts[0], gt[0])
tl = tr.step_timeline()
st = tr.stage_timings()
print(json.dumps({"in_graph_ms": {k: round(v, 4) for k, v in tl.items()}, "isolated_ms": {k: round(v, 4) for k, v in st.items()}}))
