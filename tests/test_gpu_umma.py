"""The tcgen05 deformation-MLP kernel (csrc/field_umma.cu) against the mma.sync kernel of csrc/field.cu on identical
inputs and weights: same fp16 operands and fp32 accumulation, so outputs agree to one fp16 ulp of the layer outputs
(accumulation ORDER differs between the two tensor-core paths): deform atol 2e-3 * max|dx|, saved activations identical
up to isolated 1-ulp rounding flips (checked as <= 2^-9 relative on > 99.9% of entries)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(cuda_dev, M, seed=0):
    from seald_nerf_b200 import field as F
    from seald_nerf_b200.dnerf.network import NeRFNetwork
    torch.manual_seed(seed)
    net = NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True).to(cuda_dev)
    for l in net.deform_net:
        l.weight.data.mul_(1.5)  # a deformation field that actually moves points
    cfg = net._field_cfg
    hw = F.HalfWeights(cfg, cuda_dev)
    hw.refresh([w.detach() for w in net.mlp_weights()])
    g = torch.Generator().manual_seed(seed + 1)
    xyz = (torch.rand(M, 3, generator=g) * 1.8 - 0.9).to(cuda_dev)
    return F, cfg, hw, xyz


def _run(F, cfg, hw, xyz, tval, impl, save, m_live=None, t0_mode=1):
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    M = xyz.shape[0]
    dev = xyz.device
    deform = torch.full((M, 3), 7.0, device=dev); x01 = torch.full((M, 3), 7.0, device=dev)
    Mp = (M + 127) // 128 * 128 if impl == "umma" else M  # the tcgen05 kernels save whole 128-row tile images (field.tile_image)
    in_buf = torch.zeros(Mp, 80, dtype=torch.float16, device=dev) if save else None
    fwd = torch.zeros(cfg.n_deform - 1, Mp, 128, dtype=torch.float16, device=dev) if save else None
    td = torch.tensor([tval], device=dev)
    m_dev = None if m_live is None else torch.tensor([m_live], dtype=torch.int32, device=dev)
    old = F.DEFORM_IMPL
    F.DEFORM_IMPL = impl
    try:
        F.deform_forward(cfg, hw, xyz, td, M, m_dev, t0_mode, deform, x01, in_buf, fwd)
    finally:
        F.DEFORM_IMPL = old
    torch.cuda.synchronize()
    if save and impl == "umma":  # back to row-major for the comparisons
        in_buf = F.from_tile_image(in_buf)
        fwd = torch.stack([F.from_tile_image(f) for f in fwd])
    return deform, x01, in_buf, fwd


@pytest.mark.parametrize("M,m_live", [(1000, None), (128 * 5 + 3, None), (128, None), (4096 * 11, 4096 * 7 + 77), (300, 0), (200000, None)])
def test_umma_deform_matches_mma_sync(cuda_dev, M, m_live):
    F, cfg, hw, xyz = _setup(cuda_dev, M)
    d0, x0, i0, f0 = _run(F, cfg, hw, xyz, 0.37, "mma", True, m_live)
    d1, x1, i1, f1 = _run(F, cfg, hw, xyz, 0.37, "umma", True, m_live)
    n = M if m_live is None else m_live
    if n == 0:
        assert float((d1 - 7.0).abs().max()) == 0.0  # nothing written
        return
    assert torch.equal(i0[:n], i1[:n]), "encoded inputs are computed by the same expressions"
    scale = float(d0[:n].abs().max())
    assert scale > 1e-3
    assert float((d0[:n] - d1[:n]).abs().max()) <= 2e-3 * scale + 1e-6
    torch.testing.assert_close(x1[:n], x0[:n], rtol=0, atol=2e-3 * scale + 1e-6)
    a, b = f0[:, :n].float(), f1[:, :n].float()
    rel = (a - b).abs() / (a.abs().clamp(min=1e-3))
    assert float((rel <= 2.0 ** -9).float().mean()) > 0.999
    assert float((a - b).abs().max()) <= 0.02 * float(a.abs().max())
    if m_live is not None:  # rows beyond the live count are left alone
        assert float((d1[n:] - 7.0).abs().max()) == 0.0 and float(f1[:, n:].abs().max()) == 0.0


def test_umma_deform_inference_and_t0_modes(cuda_dev):
    F, cfg, hw, xyz = _setup(cuda_dev, 777, seed=3)
    d_tr, x_tr, _, _ = _run(F, cfg, hw, xyz, 0.6, "umma", True)
    d_inf, x_inf, _, _ = _run(F, cfg, hw, xyz, 0.6, "umma", False)
    assert torch.equal(d_tr, d_inf) and torch.equal(x_tr, x_inf)
    # t == 0: forward() semantics zero the deformation, density() semantics still report it (network.py:140-141,188-190)
    d1, x1, _, _ = _run(F, cfg, hw, xyz, 0.0, "umma", False, t0_mode=1)
    d2, x2, _, _ = _run(F, cfg, hw, xyz, 0.0, "umma", False, t0_mode=2)
    assert float(d1.abs().max()) == 0.0 and float(d2.abs().max()) > 0.0
    torch.testing.assert_close(x1, (xyz + 1) / 2, rtol=0, atol=1e-7)
    assert torch.equal(x1, x2)
    d2m, _, _, _ = _run(F, cfg, hw, xyz, 0.0, "mma", False, t0_mode=2)
    assert float((d2 - d2m).abs().max()) <= 2e-3 * float(d2m.abs().max()) + 1e-6


@pytest.mark.parametrize("M,m_live", [(1000, None), (4096 * 11, 4096 * 7 + 77), (64, None), (300, 0)])
def test_umma_wgrad_matches_mma_sync(cuda_dev, M, m_live):
    """tcgen05 weight-gradient kernel (MN-major operands, csrc/wgrad_umma.cu) vs the mma.sync kernel on the same job table:
    both accumulate fp16 products in fp32; only the summation order differs -> rtol 1e-3 of the largest entry per layer."""
    from seald_nerf_b200 import field as F
    from seald_nerf_b200.dnerf.network import NeRFNetwork
    torch.manual_seed(0)
    net = NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True).to(cuda_dev)
    cfg = net._field_cfg
    ws = F.FieldWorkspace(cfg, M, cuda_dev, training=True)
    g = torch.Generator(device=cuda_dev).manual_seed(1)
    n_live = M if m_live is None else m_live
    # row-major "truth" of every operand (rows beyond the live count zero, as the kernels leave them in the saved tile images)
    tiled_bufs = ("in_buf", "fwd_d", "bwd_d", "gout_d", "cin", "fwd_s", "fwd_c", "bwd_s", "bwd_c", "gout_s", "gout_c")
    rowmajor = {}
    for name in tiled_bufs + ("feat",):
        t = getattr(ws, name)
        t.copy_(torch.randn(t.shape, device=cuda_dev, generator=g).to(t.dtype))
        t[..., n_live:, :] = 0
        rowmajor[name] = t.clone()

    def to_img(x):
        return F.tile_image(x) if x.dim() == 2 else torch.stack([F.tile_image(y) for y in x])

    m_dev = None if m_live is None else torch.tensor([m_live], dtype=torch.int32, device=cuda_dev)
    outs = []
    for impl in ("mma", "umma"):
        # mma.sync kernel: row-major operands; tcgen05 kernel: every operand as tile images (bulk-copy path), the sigma net's first
        # layer reading the tile-image copy of the features
        for name in tiled_bufs:
            getattr(ws, name).copy_(to_img(rowmajor[name]) if impl == "umma" else rowmajor[name])
        ws.heads_tiled = impl == "umma"
        if impl == "umma":
            Mp = ws.cin.shape[0]
            feat_p = torch.zeros(Mp, 32, dtype=torch.float16, device=cuda_dev)
            feat_p[:M] = rowmajor["feat"]
            ws.feat_img.copy_(F.tile_image(feat_p))
        grads = [torch.zeros_like(w, dtype=torch.float32) for w in net.mlp_weights()]
        old = F.WGRAD_IMPL, F.DEFORM_IMPL
        F.WGRAD_IMPL = F.DEFORM_IMPL = impl
        try:
            jobs, n_jobs = F.wgrad_jobs(cfg, ws, grads, deform=True)
            F.mlp_wgrad(jobs, n_jobs, M, m_dev)
            F.mlp_wgrad(jobs, n_jobs, M, m_dev)  # accumulates
        finally:
            F.WGRAD_IMPL, F.DEFORM_IMPL = old
        torch.cuda.synchronize()
        outs.append(grads)
    for a, b in zip(*outs):
        if m_live == 0:
            assert float(b.abs().max()) == 0.0
            continue
        scale = float(a.abs().max())
        assert scale > 0
        assert float((a - b).abs().max()) <= 1e-3 * scale, (tuple(a.shape), float((a - b).abs().max()), scale)


@pytest.mark.parametrize("M,m_live,tval", [(1000, None, 0.37), (4096 * 11, 4096 * 7 + 77, 0.5), (128, None, 0.9), (700, None, 0.0), (200000, None, 0.2)])
def test_umma_deform_backward_matches_mma_sync(cuda_dev, M, m_live, tval):
    """tcgen05 dgrad chain (csrc/field_umma.cu) vs the mma.sync backward on the same saved activations: activation gradients
    agree to fp16 round-off of each layer's output (accumulation order differs): <= 2^-8 relative on > 99.5% of entries
    and <= 2% of the layer maximum everywhere; dL/d(dx) (gout) bit-identical; zero at t == 0."""
    F, cfg, hw, xyz = _setup(cuda_dev, M)
    _, _, in_buf, fwd = _run(F, cfg, hw, xyz, tval, "umma", True, m_live)  # (row-major view of the saved tile images)
    Mp = (M + 127) // 128 * 128
    g = torch.Generator(device=cuda_dev).manual_seed(5)
    grad_x01 = torch.randn(M, 3, device=cuda_dev, generator=g) * 64.0
    td = torch.tensor([tval], device=cuda_dev)
    m_dev = None if m_live is None else torch.tensor([m_live], dtype=torch.int32, device=cuda_dev)
    outs = []
    for impl in ("mma", "umma"):
        rows = Mp if impl == "umma" else M
        bwd = torch.zeros(cfg.n_deform - 1, rows, 128, dtype=torch.float16, device=cuda_dev)
        gout = torch.zeros(rows, 16, dtype=torch.float16, device=cuda_dev)
        fwd_in = torch.stack([F.tile_image(f[:M]) for f in fwd]) if impl == "umma" else fwd[:, :M].contiguous()
        old = F.DEFORM_IMPL
        F.DEFORM_IMPL = impl
        try:
            F.deform_backward(cfg, hw, grad_x01, td, M, m_dev, fwd_in, bwd, gout)
        finally:
            F.DEFORM_IMPL = old
        torch.cuda.synchronize()
        if impl == "umma":
            bwd, gout = torch.stack([F.from_tile_image(b) for b in bwd]), F.from_tile_image(gout)
        outs.append((bwd, gout))
    n = M if m_live is None else m_live
    assert torch.equal(outs[0][1][:n], outs[1][1][:n])
    a, b = outs[0][0][:, :n].float(), outs[1][0][:, :n].float()
    if tval == 0.0:
        assert float(a.abs().max()) == 0.0 and float(b.abs().max()) == 0.0
        return
    assert float(a.abs().max()) > 1e-3
    for l in range(a.shape[0]):
        scale = float(a[l].abs().max())
        assert float((a[l] - b[l]).abs().max()) <= 0.02 * scale, l
        rel = (a[l] - b[l]).abs() / a[l].abs().clamp(min=1e-3 * scale)
        assert float((rel <= 2.0 ** -8).float().mean()) > 0.995, l
        assert torch.equal(a[l] == 0, b[l] == 0) or float(((a[l] == 0) != (b[l] == 0)).float().mean()) < 1e-3  # same ReLU mask
    if m_live is not None:
        assert float(outs[1][0][:, n:].abs().max()) == 0.0


def test_umma_wgrad_raises_the_overflow_flag(cuda_dev):
    """seald_mlp_wgrad_umma_flag: GradScaler's found_inf from the weight-gradient flush itself — clean operands leave the flag alone,
    one inf in an activation gradient (tile-image layout) raises it; the gradients equal those of the unflagged entry point."""
    import ctypes as C
    from seald_nerf_b200 import _lib, field as F
    from seald_nerf_b200._lib import ptr
    from seald_nerf_b200.dnerf.network import NeRFNetwork
    torch.manual_seed(0)
    net = NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True).to(cuda_dev)
    cfg = net._field_cfg
    M = 1000
    ws = F.FieldWorkspace(cfg, M, cuda_dev, training=True)
    g = torch.Generator(device=cuda_dev).manual_seed(1)
    for name in ("in_buf", "fwd_d", "bwd_d", "gout_d", "cin", "feat_img", "fwd_s", "fwd_c", "bwd_s", "bwd_c", "gout_s", "gout_c"):
        t = getattr(ws, name)
        t.copy_((torch.randn(t.shape, device=cuda_dev, generator=g) * 0.1).to(t.dtype))
    flag = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
    res = []
    for poison in (False, True):
        if poison:
            ws.bwd_d[3].view(-1)[12345] = float("inf")
        grads = [torch.zeros_like(w, dtype=torch.float32) for w in net.mlp_weights()]
        jobs, n_jobs = F.wgrad_jobs(cfg, ws, grads, deform=True)
        flag.zero_()
        _lib.call("seald_mlp_wgrad_umma_flag", C.cast(jobs, C.c_void_p), n_jobs, M, None, ptr(flag), _lib.stream())
        torch.cuda.synchronize()
        res.append((int(flag), grads))
    assert res[0][0] == 0 and res[1][0] == 0x3f800000
    plain = [torch.zeros_like(w, dtype=torch.float32) for w in net.mlp_weights()]
    ws.bwd_d[3].view(-1)[12345] = 0.0
    jobs, n_jobs = F.wgrad_jobs(cfg, ws, plain, deform=True)
    F.mlp_wgrad(jobs, n_jobs, M, None)
    torch.cuda.synchronize()
    for a, b in zip(res[0][1], plain):
        scale = float(b.abs().max())
        if scale > 0:  # (the poisoned element was random before: only that layer's gradient differs slightly)
            assert float((a - b).abs().max()) <= 5e-2 * scale
    assert not all(bool(torch.isfinite(x).all()) for x in res[1][1])


@pytest.mark.parametrize("M,tval", [(1000, 0.37), (128 * 3 + 5, 0.0), (76000, 0.81), (200000, 0.37), (700001, 0.5)])
def test_one_launch_density_matches_three_kernel_density(cuda_dev, M, tval):
    """seald_field_density_umma (deformation net -> hash-grid gather -> sigma head inside ONE tcgen05 tile pipeline) against the
    deformation kernel + fused grid/sigma kernel on the same inputs: same fp16 operands, fp32 accumulation in a different ORDER in the
    sigma head (tcgen05 vs mma.sync), so log-densities agree to fp16 rounding flips of the hidden activations; the occupancy scatter
    writes exactly sigma * scale.  Both tile-group configurations (G = 2 below 592 tiles, G = 4 above), points outside the grid,
    t == 0 (no deformation)."""
    from seald_nerf_b200 import _lib
    F, cfg, hw, xyz = _setup(cuda_dev, M, seed=5)
    from seald_nerf_b200.dnerf.network import NeRFNetwork
    torch.manual_seed(5)
    net = NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True).to(cuda_dev)
    table16 = (net.encoder.embeddings.detach() * 3e4).to(torch.float16)  # features of order 1: densities that vary over the batch
    xyz = xyz * 1.25  # some points leave [-bound, bound]: zero features
    ws = F.FieldWorkspace(cfg, M, cuda_dev, training=False)
    td = torch.tensor([tval], device=cuda_dev)
    old = F.DENSITY_IMPL
    res = {}
    try:
        for impl in ("split", "umma"):
            F.DENSITY_IMPL = impl
            ws.sigma.fill_(-7.0)
            idx = torch.randperm(M, device=cuda_dev).to(torch.int32)
            tmp = torch.full((M,), -1.0, device=cuda_dev)
            F.field_density(cfg, hw, ws, xyz, td, table16, net.encoder.offsets, scatter=(idx, 0.5, tmp), sigma_only=True)
            torch.cuda.synchronize()
            assert torch.equal(tmp[idx.long()], ws.sigma * 0.5)
            res[impl] = ws.sigma.clone()
    finally:
        F.DENSITY_IMPL = old
    a, b = res["split"], res["umma"]
    assert bool(torch.isfinite(b).all()) and float(b.min()) > 0
    la, lb = a.log(), b.log()
    assert float(la.std()) > 0.05, "the test field must vary"
    d = (la - lb).abs()
    assert float(d.max()) <= 0.03 and float(d.mean()) <= 1e-3, (float(d.max()), float(d.mean()))
    assert float((d <= 2e-3).float().mean()) > 0.98
