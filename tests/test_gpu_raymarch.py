"""GPU parity of the march / composite kernels against the CPU oracle (oracle/raymarch_oracle.c) and, when the
reference's own extension is present (oracle/_ref/_ref_raymarching.so), against the reference itself.

Bar (north_star): Morton indices, bitfield bytes, per-ray sample counts and offsets' validity are BIT-EXACT;
sample positions are bit-exact too (same fp32 ops); composited values within fp32 tolerance (rtol 1e-4, atol 1e-6:
the only difference is ex2.approx vs exp2f).
"""
import numpy as np
import pytest
import torch

from conftest import load_ref
from helpers import camera_rays, scene_bitfield

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-6


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_morton_roundtrip_all_cells(cuda_dev):
    from seald_nerf_b200 import raymarching as rm
    from oracle import raymarch as orc
    H = 128
    g = torch.arange(H, dtype=torch.int32)
    coords = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    idx = rm.morton3D(coords.to(cuda_dev))
    assert idx.dtype == torch.int32 and int(idx.max()) == H ** 3 - 1 and int(idx.min()) == 0
    assert np.array_equal(idx.cpu().numpy(), orc.morton3D(coords.numpy()))
    back = rm.morton3D_invert(idx)
    assert torch.equal(back.cpu(), coords)
    assert np.array_equal(orc.morton3D_invert(idx.cpu().numpy()), coords.numpy())
    assert torch.unique(idx).numel() == H ** 3


def test_morton_vs_reference(cuda_dev):
    ref = load_ref("raymarching")
    if ref is None:
        pytest.skip("reference extension not built")
    from seald_nerf_b200 import raymarching as rm
    coords = torch.randint(0, 128, (100003, 3), dtype=torch.int32, device=cuda_dev)
    out_ref = torch.empty(coords.shape[0], dtype=torch.int32, device=cuda_dev)
    ref.morton3D(coords, coords.shape[0], out_ref)
    assert torch.equal(rm.morton3D(coords), out_ref)
    inv_ref = torch.empty(coords.shape[0], 3, dtype=torch.int32, device=cuda_dev)
    ref.morton3D_invert(out_ref, coords.shape[0], inv_ref)
    assert torch.equal(rm.morton3D_invert(out_ref), inv_ref)


@pytest.mark.parametrize("n_cells", [128 ** 3, 8 * 1237, 8 * 4 * 3 + 8])
def test_packbits_bit_exact(cuda_dev, n_cells):
    from seald_nerf_b200 import raymarching as rm
    from oracle import raymarch as orc
    rng = np.random.default_rng(1)
    grid = rng.normal(0, 10, n_cells).astype(np.float32)
    grid[rng.integers(0, n_cells, n_cells // 7)] = -1.0  # untrained cells
    thresh = float(grid[5])  # a threshold equal to a cell value exercises the strict '>'
    bits = rm.packbits(_t(grid, cuda_dev).view(1, -1), thresh)
    assert bits.dtype == torch.uint8
    assert np.array_equal(bits.cpu().numpy(), orc.packbits(grid, thresh))
    ref = load_ref("raymarching")
    if ref is not None:
        out = torch.empty(n_cells // 8, dtype=torch.uint8, device=cuda_dev)
        ref.packbits(_t(grid, cuda_dev), n_cells // 8, thresh, out)
        assert torch.equal(bits, out)


def test_near_far_bit_exact(cuda_dev):
    from seald_nerf_b200 import raymarching as rm
    from oracle import raymarch as orc
    ro, rd = camera_rays(20000, seed=3)
    # edge cases: axis-parallel rays (1/0 = inf), rays starting inside the box, rays that miss
    rd[:50] = np.array([0, 0, -1], np.float32)
    rd[50:100] = np.array([1, 0, 0], np.float32)
    ro[100:200] = np.random.default_rng(0).uniform(-0.5, 0.5, (100, 3)).astype(np.float32)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = rm.near_far_from_aabb(_t(ro, cuda_dev), _t(rd, cuda_dev), _t(aabb, cuda_dev), 0.2)
    n_o, f_o = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    assert np.array_equal(nears.cpu().numpy(), n_o) and np.array_equal(fars.cpu().numpy(), f_o)
    assert (n_o == np.finfo(np.float32).max).any() and (n_o < 1e3).any()
    ref = load_ref("raymarching")
    if ref is not None:
        n_r = torch.empty_like(nears)
        f_r = torch.empty_like(fars)
        ref.near_far_from_aabb(_t(ro, cuda_dev), _t(rd, cuda_dev), _t(aabb, cuda_dev), ro.shape[0], 0.2, n_r, f_r)
        assert torch.equal(nears, n_r) and torch.equal(fars, f_r)


def _per_ray(rays, xyzs, deltas):
    """ray id -> (count, samples) irrespective of packing order."""
    out = {}
    for rid, off, cnt in rays:
        out[int(rid)] = (int(cnt), xyzs[off:off + cnt].copy(), deltas[off:off + cnt].copy())
    return out


@pytest.mark.parametrize("cascade,bound,dt_gamma,perturb", [(1, 1.0, 0.0, False), (1, 1.0, 0.0, True), (1, 1.0, 1 / 128, True),
                                                           (2, 2.0, 1 / 128, False), (2, 2.0, 0.0, True)])
def test_march_train_bit_exact(cuda_dev, cascade, bound, dt_gamma, perturb):
    from seald_nerf_b200 import _lib
    from seald_nerf_b200 import raymarching as rm
    from seald_nerf_b200._lib import ptr
    from oracle import raymarch as orc
    N, H, max_steps = 4096, 128, 1024
    bits, _ = scene_bitfield(cascade=cascade)
    ro, rd = camera_rays(N, seed=11, center_crop=200)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    noises = np.random.default_rng(5).random(N, dtype=np.float32) if perturb else np.zeros(N, np.float32)
    xo, do_, deo, rays_o_, cnt_o = orc.march_rays_train(ro, rd, bound, bits, cascade, H, nears, fars, noises, dt_gamma, max_steps)
    total = int(cnt_o[0])
    assert total > 10000, "scene should produce a realistic number of samples"
    M = total + 128

    d = cuda_dev
    xyzs = torch.zeros(M, 3, device=d); dirs = torch.zeros(M, 3, device=d); deltas = torch.zeros(M, 2, device=d)
    rays = torch.zeros(N, 3, dtype=torch.int32, device=d); counter = torch.zeros(2, dtype=torch.int32, device=d)
    tro, trd, tb, tn, tf, tz = _t(ro, d), _t(rd, d), _t(bits, d), _t(nears, d), _t(fars, d), _t(noises, d)
    _lib.call("seald_march_rays_train", ptr(tro), ptr(trd), ptr(tb), bound, dt_gamma, max_steps, N, cascade, H, M, ptr(tn), ptr(tf),
              None, 0.0, None, None, ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(counter), ptr(tz), None, _lib.stream())
    torch.cuda.synchronize()
    rays_c = rays.cpu().numpy()
    assert counter.cpu().tolist() == [total, N]
    # per-ray sample counts bit-exact, ray ids in order, ranges form an exact disjoint packing of [0, total)
    assert np.array_equal(rays_c[:, 0], np.arange(N))
    assert np.array_equal(rays_c[:, 2], rays_o_[:, 2])
    order = np.lexsort((rays_c[:, 2], rays_c[:, 1]))  # by offset, empty rays first among ties
    offs, cnts = rays_c[order, 1], rays_c[order, 2]
    assert offs[0] == 0 and np.array_equal(offs[1:], np.cumsum(cnts)[:-1]) and offs[-1] + cnts[-1] == total
    # samples bit-exact per ray
    mine = _per_ray(rays_c, xyzs.cpu().numpy(), deltas.cpu().numpy())
    gold = _per_ray(rays_o_, xo, deo)
    for rid in range(N):
        assert mine[rid][0] == gold[rid][0]
        assert np.array_equal(mine[rid][1], gold[rid][1]), "xyz mismatch on ray %d" % rid
        assert np.array_equal(mine[rid][2], gold[rid][2]), "delta mismatch on ray %d" % rid
    # dirs are the ray direction replicated
    dirs_c = dirs.cpu().numpy()
    rid_of_sample = np.repeat(rays_c[order, 0], cnts)
    assert np.array_equal(dirs_c[:total], rd[rid_of_sample])

    # fused AABB variant gives the same result
    xyzs2 = torch.zeros_like(xyzs); dirs2 = torch.zeros_like(dirs); deltas2 = torch.zeros_like(deltas)
    rays2 = torch.zeros_like(rays); counter2 = torch.zeros_like(counter)
    n_out = torch.empty(N, device=d); f_out = torch.empty(N, device=d)
    taabb = _t(aabb, d)
    _lib.call("seald_march_rays_train", ptr(tro), ptr(trd), ptr(tb), bound, dt_gamma, max_steps, N, cascade, H, M, None, None,
              ptr(taabb), 0.2, ptr(n_out), ptr(f_out), ptr(xyzs2), ptr(dirs2), ptr(deltas2), ptr(rays2), ptr(counter2), ptr(tz),
              ptr(rm.occupancy_aabb(tb, cascade, H, bound)), _lib.stream())  # fused slab test AND the occupied-region guard
    assert np.array_equal(n_out.cpu().numpy(), nears) and np.array_equal(f_out.cpu().numpy(), fars)
    assert torch.equal(rays2[:, 2], rays[:, 2])
    mine2 = _per_ray(rays2.cpu().numpy(), xyzs2.cpu().numpy(), deltas2.cpu().numpy())
    for rid in range(0, N, 7):
        assert np.array_equal(mine2[rid][1], gold[rid][1])

    # the reference's own kernel: identical counts and samples per ray (its packing order is atomics-dependent)
    ref = load_ref("raymarching")
    if ref is not None:
        xr = torch.zeros(M, 3, device=d); dr = torch.zeros(M, 3, device=d); der = torch.zeros(M, 2, device=d)
        rr = torch.zeros(N, 3, dtype=torch.int32, device=d); cr = torch.zeros(2, dtype=torch.int32, device=d)
        ref.march_rays_train(tro, trd, tb, bound, dt_gamma, max_steps, N, cascade, H, M, tn, tf, xr, dr, der, rr, cr, tz)
        torch.cuda.synchronize()
        assert cr.cpu().tolist() == [total, N]
        theirs = _per_ray(rr.cpu().numpy(), xr.cpu().numpy(), der.cpu().numpy())
        assert sorted(theirs) == list(range(N))
        for rid in range(N):
            assert theirs[rid][0] == mine[rid][0]
            assert np.array_equal(theirs[rid][1], mine[rid][1]), "xyz differs from the reference on ray %d" % rid
            assert np.array_equal(theirs[rid][2], mine[rid][2])


def test_march_train_overflow_and_empty(cuda_dev):
    """M too small: every kept ray's range is in bounds and disjoint; dropped rays composite to zero.
    Empty bitfield: no samples.  Full bitfield: counts hit max_steps."""
    from seald_nerf_b200 import raymarching as rm
    N, H = 2048, 128
    bits, _ = scene_bitfield()
    ro, rd = camera_rays(N, seed=2, center_crop=150)
    d = cuda_dev
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1.0], device=d)
    tro, trd = _t(ro, d), _t(rd, d)
    nears, fars = rm.near_far_from_aabb(tro, trd, aabb, 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device=d)
    xyzs, dirs, deltas, rays = rm.march_rays_train(tro, trd, 1.0, _t(bits, d), 1, H, nears, fars, counter, 4000, False, 128, False, 0, 1024)
    assert xyzs.shape[0] == 4096  # 4000 rounded up to 128 (raymarching.py:200-203)
    total = int(counter[0])
    assert total > 4096
    r = rays.cpu().numpy()
    kept = (r[:, 1] + r[:, 2] <= 4096) & (r[:, 2] > 0)
    assert kept.any() and (~kept & (r[:, 2] > 0)).any()
    sig = torch.rand(4096, device=d) * 20; rgb = torch.rand(4096, 3, device=d)
    ws, depth, image = rm.composite_rays_train(sig, rgb, deltas, rays)
    dropped = torch.from_numpy(~kept).to(d)
    assert float(ws[dropped].abs().max()) == 0 and float(image[dropped].abs().max()) == 0
    assert float(ws[~dropped].min()) > 0
    # empty grid
    counter.zero_()
    empty = torch.zeros_like(_t(bits, d))
    x2, _, _, rays2 = rm.march_rays_train(tro, trd, 1.0, empty, 1, H, nears, fars, counter, -1, False, 128, True, 0, 1024)
    assert int(counter[0]) == 0 and int(rays2[:, 2].max()) == 0 and x2.shape[0] == 128  # m += 128 - 0 % 128
    # full grid: every hitting ray takes samples until far or max_steps
    counter.zero_()
    full = torch.full_like(empty, 255)
    _, _, _, rays3 = rm.march_rays_train(tro, trd, 1.0, full, 1, H, nears, fars, counter, -1, False, 128, True, 0, 64)
    assert int(rays3[:, 2].max()) == 64


def _random_samples(rays_np, M, seed=0):
    rng = np.random.default_rng(seed)
    sig = (rng.random(M, dtype=np.float32) * 30).astype(np.float32)
    rgb = rng.random((M, 3), dtype=np.float32)
    return sig, rgb


@pytest.mark.parametrize("T_thresh", [1e-4, 1e-2])
def test_composite_train_fwd_bwd(cuda_dev, T_thresh):
    from seald_nerf_b200 import raymarching as rm
    from oracle import raymarch as orc
    N, H = 4096, 128
    bits, _ = scene_bitfield()
    ro, rd = camera_rays(N, seed=4, center_crop=200)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, 0.2)
    xo, do_, deo, rays_np, cnt = orc.march_rays_train(ro, rd, 1.0, bits, 1, H, nears, fars)
    M = int(cnt[0]) + 100
    xo, deo = xo[:M], deo[:M]
    sig, rgb = _random_samples(rays_np, M)
    d = cuda_dev
    ts = _t(sig, d).requires_grad_(True); tc = _t(rgb, d).requires_grad_(True)
    tde, tr = _t(deo, d), _t(rays_np, d)
    ws, depth, image = rm.composite_rays_train(ts, tc, tde, tr, T_thresh)
    ws_o, depth_o, image_o = orc.composite_rays_train_forward(sig, rgb, deo, rays_np, T_thresh)
    np.testing.assert_allclose(ws.detach().cpu().numpy(), ws_o, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), depth_o, rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(image.detach().cpu().numpy(), image_o, rtol=RTOL, atol=ATOL)
    gws = torch.rand(N, device=d); gim = torch.rand(N, 3, device=d)
    (ws * gws).sum().add((image * gim).sum()).add(depth.sum()).backward()  # grad_depth must be ignored
    gs_o, gc_o = orc.composite_rays_train_backward(gws.cpu().numpy(), gim.cpu().numpy(), sig, rgb, deo, rays_np, ws_o, image_o, T_thresh)
    np.testing.assert_allclose(tc.grad.cpu().numpy(), gc_o, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(ts.grad.cpu().numpy(), gs_o, rtol=1e-3, atol=2e-5)
    # early stop leaves trailing gradients at exactly zero
    assert (ts.grad == 0).sum() > 0
    ref = load_ref("raymarching")
    if ref is not None:
        ws_r = torch.empty(N, device=d); dp_r = torch.empty(N, device=d); im_r = torch.empty(N, 3, device=d)
        ref.composite_rays_train_forward(ts.detach(), tc.detach(), tde, tr, M, N, T_thresh, ws_r, dp_r, im_r)
        assert torch.equal(ws_r, ws.detach()) and torch.equal(im_r, image.detach()) and torch.equal(dp_r, depth.detach())
        gs_r = torch.zeros(M, device=d); gc_r = torch.zeros(M, 3, device=d)
        ref.composite_rays_train_backward(gws, gim, ts.detach(), tc.detach(), tde, tr, ws_r, im_r, M, N, T_thresh, gs_r, gc_r)
        assert torch.equal(gc_r, tc.grad) and torch.equal(gs_r, ts.grad)


@pytest.mark.parametrize("T_thresh", [1e-2, 1e-4])
def test_inference_loop_matches_oracle(cuda_dev, T_thresh):
    """The whole march_rays / composite_rays loop of run_cuda's eval branch (dnerf/renderer.py:350-376) on 20k rays,
    with a synthetic field (sigma, rgb from position) evaluated identically on both sides."""
    from seald_nerf_b200 import raymarching as rm
    from oracle import raymarch as orc
    N, H, max_steps = 20000, 128, 1024
    bits, _ = scene_bitfield()
    ro, rd = camera_rays(N, seed=9, center_crop=260)
    aabb = np.array([-1, -1, -1, 1, 1, 1], np.float32)
    nears, fars = orc.near_far_from_aabb(ro, rd, aabb, 0.2)

    def field_np(x):
        sig = (np.abs(np.sin(x[:, 0] * 9) * 40)).astype(np.float32)
        rgb = (0.5 + 0.5 * np.cos(x * 5)).astype(np.float32)
        return sig, rgb

    d = cuda_dev
    tro, trd, tb, tn, tf = _t(ro, d), _t(rd, d), _t(bits, d), _t(nears, d), _t(fars, d)
    ws = torch.zeros(N, device=d); depth = torch.zeros(N, device=d); image = torch.zeros(N, 3, device=d)
    alive = torch.arange(N, dtype=torch.int32, device=d); rays_t = tn.clone()
    ws_o = np.zeros(N, np.float32); depth_o = np.zeros(N, np.float32); image_o = np.zeros((N, 3), np.float32)
    alive_o = np.arange(N, dtype=np.int32); rays_t_o = nears.copy()
    ref = load_ref("raymarching")
    step = 0
    n_iter = 0
    while step < max_steps:
        n_alive = alive.shape[0]
        assert n_alive == alive_o.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = rm.march_rays(n_alive, n_step, alive, rays_t, tro, trd, 1.0, tb, 1, H, tn, tf, 128, False, 0, max_steps)
        xo, do_, deo = orc.march_rays(n_alive, n_step, alive_o, rays_t_o, ro, rd, 1.0, bits, 1, H, nears, fars, 128)
        assert xyzs.shape[0] == xo.shape[0]
        assert np.array_equal(xyzs.cpu().numpy(), xo) and np.array_equal(deltas.cpu().numpy(), deo) and np.array_equal(dirs.cpu().numpy(), do_)
        if ref is not None and n_iter < 3:
            xr = torch.zeros_like(xyzs); dr = torch.zeros_like(dirs); der = torch.zeros_like(deltas)
            ref.march_rays(n_alive, n_step, alive, rays_t, tro, trd, 1.0, 0.0, max_steps, 1, H, tb, tn, tf, xr, dr, der,
                           torch.zeros(n_alive, device=d))
            assert torch.equal(xr, xyzs) and torch.equal(der, deltas)
        sig, rgb = field_np(xo)
        rm.composite_rays(n_alive, n_step, alive, rays_t, _t(sig, d), _t(rgb, d), deltas, ws, depth, image, T_thresh)
        orc.composite_rays(n_alive, n_step, alive_o, rays_t_o, sig, rgb, deo, ws_o, depth_o, image_o, T_thresh)
        assert np.array_equal(alive.cpu().numpy(), alive_o), "alive/dead decisions must match"
        # device-side compaction == boolean mask
        comp, n_out = rm.compact_alive(alive)
        alive = alive[alive >= 0]
        assert int(n_out) == alive.shape[0] and torch.equal(comp[:alive.shape[0]], alive)
        alive_o = alive_o[alive_o >= 0]
        step += n_step
        n_iter += 1
    assert n_iter > 5
    np.testing.assert_allclose(ws.cpu().numpy(), ws_o, rtol=RTOL, atol=2e-6)
    np.testing.assert_allclose(image.cpu().numpy(), image_o, rtol=RTOL, atol=2e-6)
    np.testing.assert_allclose(depth.cpu().numpy(), depth_o, rtol=RTOL, atol=2e-5)
    assert float(ws.max()) > 0.9


@pytest.mark.parametrize("n_rays,dt_gamma,cascade", [(3000, 0.0, 1), (3000, 1.0 / 128, 2), (70000, 0.0, 1)])
def test_occupied_region_guard_changes_nothing(cuda_dev, n_rays, dt_gamma, cascade):
    """occ_aabb6 (seald_occupancy_aabb) lets rays skip the empty volume; counts, positions and deltas must be bit-identical
    with and without it - for the warp-per-ray kernel, the thread-per-ray kernel (> 65536 rays) and the inference march."""
    from seald_nerf_b200 import _lib
    from seald_nerf_b200 import raymarching as rm
    from seald_nerf_b200._lib import ptr
    d = cuda_dev
    bound = float(2 ** (cascade - 1))
    H = 128
    bits, _ = scene_bitfield(H=H, cascade=cascade)
    ro, rd = camera_rays(n_rays, seed=9, radius=3.2 * bound)
    tro, trd, tb = _t(ro, d), _t(rd, d), _t(bits, d)
    aabb = torch.tensor([-bound] * 3 + [bound] * 3, dtype=torch.float32, device=d)
    nears, fars = rm.near_far_from_aabb(tro, trd, aabb, 0.2)
    occ = rm.occupancy_aabb(tb, cascade, H, bound, 2)
    o = occ.cpu().numpy()
    assert np.all(o[:3] < o[3:]) and np.all(o[:3] >= -bound - 0.1) and np.all(o[3:] <= bound + 0.1)
    assert (o[3:] - o[:3]).max() < 2 * bound * 0.95, "the figure does not fill the volume: the guard must be a real restriction"
    noises = torch.rand(n_rays, device=d)
    res = []
    for use in (None, occ):
        M = n_rays * 64
        xyzs = torch.zeros(M, 3, device=d); dirs = torch.zeros(M, 3, device=d); deltas = torch.zeros(M, 2, device=d)
        rays = torch.zeros(n_rays, 3, dtype=torch.int32, device=d); counter = torch.zeros(2, dtype=torch.int32, device=d)
        _lib.call("seald_march_rays_train", ptr(tro), ptr(trd), ptr(tb), bound, dt_gamma, 1024, n_rays, cascade, H, M, ptr(nears), ptr(fars),
                  None, 0.0, None, None, ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(counter), ptr(noises), ptr(use), _lib.stream())
        torch.cuda.synchronize()
        r = rays.cpu().numpy()
        x, de = xyzs.cpu().numpy(), deltas.cpu().numpy()
        # CTAs race for output ranges: compare per ray
        res.append((counter.cpu().tolist(), r[:, 2].copy(), [x[o_:o_ + k].copy() for _, o_, k in r[::13]], [de[o_:o_ + k].copy() for _, o_, k in r[::13]]))
    assert res[0][0] == res[1][0] and res[0][0][0] > 100
    assert np.array_equal(res[0][1], res[1][1])
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)
    for a, b in zip(res[0][3], res[1][3]):
        assert np.array_equal(a, b)

    # inference march, three rounds of 8 steps with compositing in between (rays_t advances)
    outs = []
    for use in (None, occ):
        alive = torch.arange(n_rays, dtype=torch.int32, device=d)
        rays_t = nears.clone()
        ws = torch.zeros(n_rays, device=d); depth = torch.zeros(n_rays, device=d); image = torch.zeros(n_rays, 3, device=d)
        log = []
        for _ in range(3):
            n_alive = alive.shape[0]
            M = n_alive * 8 + 128
            xyzs = torch.empty(M, 3, device=d).fill_(7); dirs = torch.empty(M, 3, device=d).fill_(7); deltas = torch.empty(M, 2, device=d).fill_(7)
            _lib.call("seald_march_rays", n_alive, 8, ptr(alive), ptr(rays_t), ptr(tro), ptr(trd), bound, dt_gamma, 1024, cascade, H, ptr(tb),
                      ptr(nears), ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), None, None, None, ptr(use), _lib.stream())
            log.append((xyzs[:n_alive * 8].clone(), deltas[:n_alive * 8].clone()))
            sig = torch.full((M,), 3.0, device=d); rgb = torch.full((M, 3), 0.5, device=d)
            rm.composite_rays(n_alive, 8, alive, rays_t, sig, rgb, deltas, ws, depth, image, 1e-4)
            alive = alive[alive >= 0]
        outs.append((log, alive.clone(), rays_t.clone(), image.clone()))
    for (xa, da), (xb, db) in zip(outs[0][0], outs[1][0]):
        assert torch.equal(xa, xb) and torch.equal(da, db)
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2]) and torch.equal(outs[0][3], outs[1][3])
