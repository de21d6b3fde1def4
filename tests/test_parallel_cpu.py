"""world_size-2 `gloo` tests (CPU) of the multi-GPU plumbing in seald_nerf_b200/parallel.py: ray-tile sharding + frame
gather, flat-gradient all-reduce with the global-mean loss normalisation, sample-buffer size agreement."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from seald_nerf_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn, world=2):
    ret = mp.Manager().dict()
    mp.spawn(_run, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return dict(ret)


def test_shard_tiles_partition_the_frame():
    for n, w, tile in ((640000, 8, 256), (1000, 3, 64), (5, 2, 256), (256 * 7 + 13, 4, 256)):
        seen = torch.zeros(n, dtype=torch.int32)
        for r in range(w):
            idx = parallel.shard_tiles(n, w, r, tile)
            assert idx.shape[0] <= parallel.shard_capacity(n, w, tile)
            assert torch.all(idx[1:] > idx[:-1]) if idx.numel() > 1 else True
            seen[idx] += 1
        assert torch.all(seen == 1), "every ray belongs to exactly one rank"


def _frame_job(rank, world):
    n = 256 * 5 + 77  # ragged: the last tile is partial and the shards have different sizes
    g = torch.Generator().manual_seed(0)
    full = torch.rand(n, 5, generator=g)          # what a single-GPU render would produce (same on every rank)
    idx = parallel.shard_tiles(n, world, rank, 256)
    local = full[idx]                              # this rank "renders" only its tiles
    out = parallel.gather_frame(local, n, rank, world, None, 256)
    return bool(torch.equal(out, full))


def test_gather_frame_gloo_world2():
    res = _spawn(_frame_job, 2)
    assert res == {0: True, 1: True}


def _grad_job(rank, world):
    # toy "field": prediction = rays @ W; every rank has its own batch; loss = sum((pred-gt)^2) * inv_count(global)
    g = torch.Generator().manual_seed(1)
    W = torch.randn(3, 3, generator=g)
    rays = torch.randn(world, 64, 3, generator=g)
    gt = torch.randn(world, 64, 3, generator=g)
    inv = parallel.loss_inv_count(64, world)
    Wl = W.clone().requires_grad_(True)
    (((rays[rank] @ Wl - gt[rank]) ** 2).sum() * inv).backward()
    flat = Wl.grad.reshape(-1).clone()
    parallel.allreduce_flat_grads(flat)
    # single-process gradient of the mean loss over the GLOBAL batch
    Wg = W.clone().requires_grad_(True)
    ((rays.reshape(-1, 3) @ Wg - gt.reshape(-1, 3)) ** 2).mean().backward()
    m = parallel.agree_max(100 + 7 * rank, torch.device("cpu"))
    return bool(torch.allclose(flat, Wg.grad.reshape(-1), rtol=1e-5, atol=1e-6)), m


def test_flat_grad_allreduce_equals_global_batch_gradient_gloo_world2():
    res = _spawn(_grad_job, 2)
    assert res[0] == (True, 107) and res[1] == (True, 107)


def test_single_process_is_a_noop():
    t = torch.ones(4)
    assert parallel.allreduce_flat_grads(t) is t and parallel.agree_max(5, torch.device("cpu")) == 5
    assert parallel.world() == (0, 1)
    x = torch.rand(10, 5)
    assert parallel.gather_frame(x, 10, 0, 1) is x


def test_shard_layout_covers_the_table_in_equal_aligned_shards():
    """Sharded optimiser / fused exchange: every element of the table belongs to exactly one rank, shards are equal, 8-element
    aligned (16-byte fp16 vectors), and the padding is smaller than 8 elements per rank."""
    from seald_nerf_b200 import parallel
    n_table = 12239728  # L16 / T2^19 / F2 hash grid
    for W in (1, 2, 3, 4, 8, 16):
        shard, padded = parallel.shard_layout(n_table, W)
        if W == 1:
            assert shard == n_table and padded % 4 == 0 and 0 <= padded - n_table < 4
            continue
        assert shard % 8 == 0 and padded == shard * W and padded >= n_table and padded - n_table < 8 * W + W
        owners = [(r * shard, min((r + 1) * shard, n_table)) for r in range(W)]
        assert owners[0][0] == 0 and owners[-1][1] == n_table
        assert all(owners[r][1] == owners[r + 1][0] or owners[r + 1][0] >= n_table for r in range(W - 1))
    for n in (1, 7, 8, 9, 1000003):
        for W in (2, 8):
            shard, padded = parallel.shard_layout(n, W)
            assert shard % 8 == 0 and padded >= n and padded == shard * W
