"""N-GPU microbenchmark of the fused exchange+optimiser kernel (csrc/dp_fused.cu): all ranks launch it together, device-timed."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from seald_nerf_b200.trainer import FusedTrainer  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
for mc, mcr in (("1", "1"), ("1", "0"), ("0", "0")):
    os.environ["SEALD_DP_MULTICAST"] = mc
    os.environ["SEALD_DP_MULTICAST_REDUCE"] = mcr
    model = bench.build_scene(dev, seed=0)
    tr = FusedTrainer(model, num_rays=4096, max_samples=4096 * 12, world_size=world, dp_mode="fused", use_graph=False)
    tr.grads.normal_()
    tr.grads[tr.n_flag:].zero_()
    from seald_nerf_b200 import _lib
    import ctypes as C

    def reduce():
        _lib.call("seald_dp_reduce_shard", C.cast(tr._peer_grads, C.c_void_p), tr._mc_grads if tr._mc_reduce else None, world,
                  tr.rank * tr.shard_len, tr.shard_len, tr.grad_shard.data_ptr(), _lib.stream())

    res = []
    for fn in (reduce, tr._optimizer):
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        tr._symm[0].barrier(0)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res.append(float(t))
    if rank == 0:
        mb = tr.shard_len * 4 / 1e6
        print("world %d multicast st %s ld_reduce %s blocks %s: reduce_shard %.4f ms (%.0f GB/s inbound), adam_broadcast stage (+fp16 weight refresh, loss scale) %.4f ms"
              % (world, mc, mcr, os.environ.get("SEALD_DP_BLOCKS", "default"), res[0], mb * (world - 1) / world / res[0], res[1]), flush=True)
    dist.barrier()
    del tr, model
dist.barrier()
os._exit(0)
