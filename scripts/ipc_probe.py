"""2-GPU probe: plain CUDA IPC (cudaIpcGetMemHandle through torch's tensor reductions) for peer-mapped buffers."""
import os
import torch
import torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
t = torch.full((1 << 20,), float(rank + 1), device=dev)
fn, args = reduce_tensor(t)
objs = [None] * world
dist.all_gather_object(objs, (fn, args))
peers = []
for r in range(world):
    peers.append(t if r == rank else objs[r][0](*objs[r][1]))
torch.cuda.synchronize()
dist.barrier()
p = peers[(rank + 1) % world]
print(rank, "peer tensor device", p.device, "ptr", hex(p.data_ptr()), flush=True)
# read the peer's memory with a kernel running on THIS device: torch ops dispatch on the tensor's device, so use a copy kernel
# launched explicitly here: index_select on local device of a peer pointer is not expressible in torch -> use our library
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seald_nerf_b200 import _lib
out = torch.zeros(4, device=dev)
can = torch.cuda.can_device_access_peer(dev.index, p.device.index)
print(rank, "can_device_access_peer", can, flush=True)
local_copy = torch.empty(1 << 20, device=dev)
local_copy.copy_(p)   # P2P copy through torch
torch.cuda.synchronize()
print(rank, "peer value via copy", float(local_copy[0]), flush=True)
dist.barrier()
os._exit(0)
