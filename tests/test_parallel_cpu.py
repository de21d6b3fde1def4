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
