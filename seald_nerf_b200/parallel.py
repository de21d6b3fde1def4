"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` (NCCL on the GPUs; the same
functions run under gloo on CPU tensors, which is how tests/test_parallel_cpu.py covers them).

Rays and pixels are independent units and the model (23-47 MB grid + 0.12 M MLP weights + 16 MiB bitfield) is replicated:
  * training  : every rank draws its own ray batch; ONE exchange per step — all-reduce(SUM) of the flat fp32 gradient
                buffer [grid table | MLP weights]; each rank's loss is already divided by the GLOBAL element count, so the
                sum is the gradient of the global-batch mean loss.  The sample-buffer size is agreed with an all-reduce(MAX).
  * rendering : interleaved ray tiles per rank, no exchange during the march loop, one all-gather of 5 floats per ray.
"""
import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_flat_grads(flat, group=None):
    """In-place SUM of the flat gradient buffer over the ranks (no-op for a single process)."""
    if world(group)[1] > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def agree_max(value, device, group=None):
    """max over ranks of a host integer (sample-buffer size M, so every rank captures the same graph shape)."""
    if world(group)[1] == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def loss_inv_count(n_rays_local, world_size, channels=3):
    """1 / (elements of the GLOBAL batch): the per-rank MSE sums scaled by this add up to the global mean loss."""
    return 1.0 / (channels * n_rays_local * world_size)


def shard_layout(n_table, world_size):
    """Layout of the flat table region for a sharded optimiser: (shard_len, n_table_pad).  The table is padded to `world_size` equal
    shards whose length is a multiple of 8 elements (16-byte fp16 vectors / two float4 per thread in the exchange kernels); rank r owns
    elements [r * shard_len, (r + 1) * shard_len).  A single rank keeps the table as it is, rounded up to a float4."""
    n_table, world_size = int(n_table), int(world_size)
    if world_size <= 1:
        return n_table, (n_table + 3) // 4 * 4
    shard_len = ((n_table + world_size - 1) // world_size + 7) // 8 * 8
    return shard_len, shard_len * world_size


def shard_tiles(n_rays, world_size, rank, tile=256):
    """Indices (int64, ascending) of the rays rank `rank` renders: tiles of `tile` consecutive rays dealt round-robin."""
    n_tiles = (n_rays + tile - 1) // tile
    mine = torch.arange(rank, n_tiles, world_size)
    idx = (mine[:, None] * tile + torch.arange(tile)[None, :]).reshape(-1)
    return idx[idx < n_rays]


def shard_capacity(n_rays, world_size, tile=256):
    """Rays of the largest shard (every rank pads its result to this many rows for the all-gather)."""
    n_tiles = (n_rays + tile - 1) // tile
    return ((n_tiles + world_size - 1) // world_size) * tile


_UNSHARD_CACHE = {}


def unshard_index(n_rays, world_size, tile, device):
    """Row of the padded, rank-major gather buffer [world * cap] that holds ray i (cached per shape and device)."""
    key = (n_rays, world_size, tile, str(device))
    idx = _UNSHARD_CACHE.get(key)
    if idx is None:
        cap = shard_capacity(n_rays, world_size, tile)
        idx = torch.empty(n_rays, dtype=torch.int64)
        for r in range(world_size):
            mine = shard_tiles(n_rays, world_size, r, tile)
            idx[mine] = r * cap + torch.arange(mine.shape[0])
        idx = idx.to(device)
        _UNSHARD_CACHE[key] = idx
    return idx


def unshard(gathered, n_rays, world_size, tile=256):
    """gathered [world, cap, K] (rank-major, padded) -> [n_rays, K] in ray order (one gather)."""
    flat = gathered.reshape(-1, gathered.shape[-1])
    return flat.index_select(0, unshard_index(n_rays, world_size, tile, gathered.device))


def gather_frame(local, n_rays, rank, world_size, group=None, tile=256):
    """local [n_local, K] (this rank's rays in shard order) -> [n_rays, K] on every rank."""
    if world_size == 1:
        return local
    cap = shard_capacity(n_rays, world_size, tile)
    packed = torch.zeros(cap, local.shape[1], dtype=local.dtype, device=local.device)
    packed[:local.shape[0]] = local
    gathered = torch.empty(world_size * cap, local.shape[1], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, packed, group=group)
    return unshard(gathered.view(world_size, cap, -1), n_rays, world_size, tile)
