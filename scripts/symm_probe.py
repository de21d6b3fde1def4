"""2-GPU probe: does torch symmetric memory (peer-mapped buffers, signal pads, multicast) work on this box?"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
t.fill_(rank + 1)
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "ptrs", [hex(p) for p in h.buffer_ptrs], "signal", [hex(p) for p in h.signal_pad_ptrs], "pad size", h.signal_pad_size,
      "multicast_ptr", hex(h.multicast_ptr or 0), flush=True)
h.barrier(0)
peer = h.get_buffer((rank + 1) % world, (1 << 20,), torch.float32)
print(rank, "peer value", float(peer[0]), float(peer[-1]), flush=True)
# graph capture of the barrier
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    h.barrier(0)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
try:
    with torch.cuda.graph(g):
        h.barrier(0)
        t.add_(1)
        h.barrier(1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    print(rank, "graph barrier ok", float(t[0]), flush=True)
except Exception as e:
    print(rank, "graph barrier failed", repr(e), flush=True)
dist.barrier()
os._exit(0)
