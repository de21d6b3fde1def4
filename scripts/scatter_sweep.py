"""Sweep of the table-scatter launch parameters on the training-step workload (stage timing of FusedTrainer)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys, torch
sys.path.insert(0, %r)
import bench
from seald_nerf_b200.trainer import FusedTrainer
dev = torch.device("cuda", 0)
model = bench.build_scene(dev)
ro, rd, ts, gt = bench.make_batches(2, dev, 0)
tr = FusedTrainer(model, num_rays=4096, max_samples=42368, use_graph=False)
tr.set_inputs(ro[0], rd[0], ts[0], gt[0])
st = tr.stage_timings(reps=20)
print("RESULT", os.environ.get("SEALD_GRID_AGG_LEVELS"), os.environ.get("SEALD_GRID_SCATTER_PPC"), round(st["grid_scatter"], 4), st["live_samples"])
''' % ROOT
for agg in ("0", "2", "4", "6", "9", "16"):
    for ppc in ("256", "1024"):
        env = dict(os.environ, SEALD_GRID_AGG_LEVELS=agg, SEALD_GRID_SCATTER_PPC=ppc)
        p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print([l for l in p.stdout.splitlines() if l.startswith("RESULT")] or p.stderr[-500:], flush=True)
