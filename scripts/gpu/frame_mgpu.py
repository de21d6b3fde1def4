"""800x800 frame render sharded over the ranks for several round schedules (measurement script, run under torchrun)."""
import os, sys, json, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from seald_nerf_b200 import microbench
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
model = microbench.build_scene(dev)
for mult, ns in ((1, 8), (1, 32), (2, 32), (4, 32), (8, 32)):
    os.environ["SEALD_RENDER_SLOTS_MULT"] = str(mult); os.environ["SEALD_RENDER_MAX_NSTEP"] = str(ns)
    r = microbench.frame_render(dev, model=model, times=(0.5,), reps=5, rank=rank, world_size=world)
    t = torch.tensor([r["t=0.50"]["ms"]], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(world, "gpus  slots x%d  max_n_step %d:" % (mult, ns), round(float(t), 3), "ms  rounds", r["t=0.50"]["iterations"], "samples(rank0)", r["t=0.50"]["field_samples"], flush=True)
dist.barrier()
dist.destroy_process_group()
