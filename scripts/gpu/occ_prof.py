"""torch.profiler timeline of one occupancy refresh on a trained model (measurement script)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from seald_nerf_b200.trainer import FusedTrainer
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
model = bench.build_scene(dev)
tr = FusedTrainer(model, num_rays=4096, max_samples=42368, lr=1e-2, lr_net=1e-3)
mode = sys.argv[1]; steps = int(sys.argv[2])
ro, rd, ts, gt = bench.make_batches(4, dev, 0)
for i in range(steps):
    tr.train_step(ro[i % 4], rd[i % 4], ts[i % 4], gt[i % 4])
tr.flush(); torch.cuda.synchronize()
it = 0 if mode == "full" else 16
model.iter_density = it; tr.update_extra_state(); torch.cuda.synchronize()
model.iter_density = it
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter(); tr.update_extra_state(); torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
print(mode, "steps", steps, "wall ms", round(wall, 2))
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
ev.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev)
print("sum of device time ms", round(tot / 1e3, 2))
for e in ev[:14]:
    print("%8.2f ms %5d x  %s" % (e.device_time_total / 1e3, e.count, e.key[:90]))
