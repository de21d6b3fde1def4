"""One occupancy refresh (full sweep, then a partial pass) of the benchmark scene through FusedTrainer.update_extra_state
(measurement script; run under `ncu --metrics gpu__time_duration.sum` for a per-kernel launch list)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from seald_nerf_b200.trainer import FusedTrainer
dev = torch.device("cuda:0")
model = bench.build_scene(dev)
tr = FusedTrainer(model, num_rays=4096, max_samples=42368, lr=1e-2, lr_net=1e-3)
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
train_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if train_steps:
    ro, rd, ts, gt = bench.make_batches(4, dev, 0)
    for i in range(train_steps):
        tr.train_step(ro[i % 4], rd[i % 4], ts[i % 4], gt[i % 4])
    tr.flush(); torch.cuda.synchronize()
    print("trained", train_steps, "steps, loss", float(tr.loss), flush=True)
model.iter_density = 0 if mode == "full" else 16
for r in range(reps):
    if mode == "full":
        model.iter_density = 0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tr.update_extra_state()
    torch.cuda.synchronize()
    print(mode, "refresh wall ms", round((time.perf_counter() - t0) * 1e3, 2), flush=True)
