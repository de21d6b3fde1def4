"""GPU parity of the Seal proxy mapping (csrc/seal.cu, fused march variants in csrc/raymarch.cu) against
 (a) the golden outputs of the REFERENCE's own seal_utils.py (tests/golden/seal.npz) and (b) the numpy oracle.

Bar: map masks bit-exact; mapped points / dirs rtol 1e-5, atol 1e-6 (brush 'linear' atol 2e-5: the reference's torch.cdist
expansion); colours atol 1e-5 (batch mean of V accumulated with float atomics).
"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import seal_cases  # noqa: E402
from helpers import camera_rays, scene_bitfield, seal_mapper_from_dict  # noqa: E402

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(HERE, "golden", "seal.npz"))
CASES = seal_cases.cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_map_to_origin_and_color_vs_reference_golden(cuda_dev, name):
    from oracle import seal as S
    mp, pts, dirs, cols = CASES[name]
    mapper = seal_mapper_from_dict(mp)
    p, d, m = mapper.map_to_origin(torch.from_numpy(pts).to(cuda_dev), torch.from_numpy(dirs).to(cuda_dev))
    assert m.dtype == torch.bool
    m_np = m.cpu().numpy()
    assert np.array_equal(m_np, GOLD[name + "_mask"]), "map mask must be bit-exact"
    atol = 2e-5 if mp["type"] == "brush" else 1e-6
    np.testing.assert_allclose(p.cpu().numpy(), GOLD[name + "_points"], rtol=1e-5, atol=atol)
    np.testing.assert_allclose(d.cpu().numpy(), GOLD[name + "_dirs"], rtol=1e-5, atol=1e-6)
    po, do_, mo = S.map_to_origin(mp, pts, dirs)
    assert np.array_equal(m_np, mo)
    np.testing.assert_allclose(p.cpu().numpy(), po, rtol=1e-5, atol=atol)
    # map_mask alone
    if mp["type"] != "anchor":
        assert torch.equal(mapper.map_mask(torch.from_numpy(pts).to(cuda_dev)), m)
    if name + "_colors" in GOLD:
        c = mapper.map_color(p[m], d[m], torch.from_numpy(cols).to(cuda_dev)[m])
        np.testing.assert_allclose(c.cpu().numpy(), GOLD[name + "_colors"], rtol=0, atol=1e-5)
        # in-place masked variant used by the renderer
        full = torch.from_numpy(cols).to(cuda_dev).clone()
        mapper.map_color_masked_(p, m, full)
        np.testing.assert_allclose(full[m].cpu().numpy(), GOLD[name + "_colors"], rtol=0, atol=1e-5)
        assert torch.equal(full[~m], torch.from_numpy(cols).to(cuda_dev)[~m])


def _scene(cuda_dev, n=2048):
    bits, _ = scene_bitfield()
    ro, rd = camera_rays(n, seed=5, center_crop=200)
    from seald_nerf_b200 import raymarching as rm
    tro, trd = torch.from_numpy(ro).to(cuda_dev), torch.from_numpy(rd).to(cuda_dev)
    aabb = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=torch.float32, device=cuda_dev)
    nears, fars = rm.near_far_from_aabb(tro, trd, aabb, 0.2)
    return torch.from_numpy(bits).to(cuda_dev), tro, trd, nears, fars


def _scene_mapper(kind):
    from oracle import seal as S
    if kind == "bbox":
        # a box around the figure's torso, moved sideways and rotated
        return S.make_bbox_mapper(center=(0.0, 0.1, 0.0), half=(0.2, 0.25, 0.2), translate=(0.08, 0.0, 0.03), rot_deg=30.0, hsv=(0.1, 0.0, 0.0))
    return S.make_brush_mapper(mode="linear", pressure=0.05, depth=0.6, attenuation=0.03, rgb=(1.0, 0.0, 0.0))


@pytest.mark.parametrize("kind", ["bbox", "brush"])
def test_fused_march_equals_march_then_map(cuda_dev, kind):
    from seald_nerf_b200 import raymarching as rm
    bits, ro, rd, nears, fars = _scene(cuda_dev)
    mapper = seal_mapper_from_dict(_scene_mapper(kind))
    N = ro.shape[0]
    # inference march, 128 steps per ray (through the whole torso)
    alive = torch.arange(N, dtype=torch.int32, device=cuda_dev)
    rays_t = nears.clone()
    x0, d0, de0 = rm.march_rays(N, 128, alive, rays_t, ro, rd, 1.0, bits, 1, 128, nears, fars, 128, False, 0, 1024)
    x1, d1, de1, m1 = rm.march_rays_seal(N, 128, alive, rays_t, ro, rd, 1.0, bits, 1, 128, nears, fars, mapper, 128, False, 0, 1024)
    xm, dm, mm = mapper.map_to_origin(x0, d0)
    assert torch.equal(de0, de1)
    assert torch.equal(m1, mm) and int(m1.sum()) > 50
    torch.testing.assert_close(x1, xm, rtol=0, atol=1e-6)
    torch.testing.assert_close(d1, dm, rtol=0, atol=1e-6)
    assert torch.equal(x1[~m1], x0[~m1])  # unmapped samples are untouched
    # training march (warp-per-ray kernel at this size)
    c0 = torch.zeros(2, dtype=torch.int32, device=cuda_dev)
    c1 = torch.zeros(2, dtype=torch.int32, device=cuda_dev)
    tx0, td0, tde0, rays0 = rm.march_rays_train(ro, rd, 1.0, bits, 1, 128, nears, fars, c0, -1, False, 128, True, 0, 1024)
    tx1, td1, tde1, rays1, tm1 = rm.march_rays_train_seal(ro, rd, 1.0, bits, 1, 128, nears, fars, mapper, c1, -1, False, 128, True, 0, 1024)
    assert torch.equal(c0, c1) and torch.equal(rays0[:, 2], rays1[:, 2])
    # both kernels pack rays of one CTA in ray order but CTAs race for ranges: compare per ray
    r0, r1 = rays0.cpu().numpy(), rays1.cpu().numpy()
    xm, dm, mm = mapper.map_to_origin(tx0, td0)
    xm, mm, tx1c, tm1c = xm.cpu().numpy(), mm.cpu().numpy(), tx1.cpu().numpy(), tm1.cpu().numpy()
    for (_, o0, k), (_, o1, _) in zip(r0[::37], r1[::37]):
        np.testing.assert_allclose(tx1c[o1:o1 + k], xm[o0:o0 + k], rtol=0, atol=1e-6)
        assert np.array_equal(tm1c[o1:o1 + k], mm[o0:o0 + k])


def test_fused_march_large_batch_thread_per_ray(cuda_dev):
    """> 65536 rays take the thread-per-ray training kernel: same samples as the warp kernel, mapped."""
    from seald_nerf_b200 import raymarching as rm
    bits, ro, rd, nears, fars = _scene(cuda_dev, n=70000)
    mapper = seal_mapper_from_dict(_scene_mapper("bbox"))
    c1 = torch.zeros(2, dtype=torch.int32, device=cuda_dev)
    c0 = torch.zeros(2, dtype=torch.int32, device=cuda_dev)
    tx0, td0, _, rays0 = rm.march_rays_train(ro, rd, 1.0, bits, 1, 128, nears, fars, c0, -1, False, 128, True, 0, 1024)
    tx1, td1, _, rays1, tm1 = rm.march_rays_train_seal(ro, rd, 1.0, bits, 1, 128, nears, fars, mapper, c1, -1, False, 128, True, 0, 1024)
    assert torch.equal(c0, c1) and torch.equal(rays0[:, 2], rays1[:, 2])
    xm, dm, mm = mapper.map_to_origin(tx0, td0)
    assert int(tm1.sum()) == int(mm.sum()) > 100


@pytest.mark.parametrize("kind", ["bbox", "brush"])
def test_teacher_render_with_mapper(cuda_dev, kind):
    """SealNeRFTeacherRenderer.run_cuda (fused mapping) == the reference's loop spelled out with the un-fused ops."""
    from seald_nerf_b200 import raymarching as rm
    from seald_nerf_b200.SealDNeRF.network import NeRFNetwork
    torch.manual_seed(0)
    net = NeRFNetwork(encoding="hashgrid", bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10).to(cuda_dev)
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    net.eval()
    bits, ro, rd, nears, fars = _scene(cuda_dev, n=1500)
    net.density_bitfield[:] = bits[None]
    mapper = seal_mapper_from_dict(_scene_mapper(kind))
    net.init_mapper(mapper=mapper)
    time = torch.tensor([[0.3]], device=cuda_dev)
    with torch.no_grad():
        out = net.render(ro[None], rd[None], time, bg_color=None, perturb=False, force_all_rays=True, dt_gamma=0, max_steps=1024)
    assert out["image"].shape == (1, 1500, 3) and out["depth"].shape == (1, 1500) and "weights_sum" in out

    # spelled-out loop (SealDNeRF/renderer.py:214-286) with march_rays -> map_to_origin -> field -> map_color -> composite_rays
    N = ro.shape[0]
    ws = torch.zeros(N, device=cuda_dev); depth = torch.zeros(N, device=cuda_dev); image = torch.zeros(N, 3, device=cuda_dev)
    alive = torch.arange(N, dtype=torch.int32, device=cuda_dev)
    rays_t = nears.clone()
    step, n_mapped = 0, 0
    with torch.no_grad():
        while step < 1024:
            n_alive = alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = rm.march_rays(n_alive, n_step, alive, rays_t, ro, rd, 1.0, bits, 1, 128, nears, fars, 128, False, 0, 1024)
            mx, md, mask = mapper.map_to_origin(xyzs, dirs)
            n_mapped += int(mask.sum())
            sig, rgb, _ = net(mx, md, time)
            rgb = rgb.float()
            rgb[mask] = mapper.map_color(mx[mask], md[mask], rgb[mask])
            rm.composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, deltas, ws, depth, image, 1e-4)
            alive = alive[alive >= 0]
            step += n_step
    image = image + (1 - ws).unsqueeze(-1)
    assert n_mapped > 100
    torch.testing.assert_close(out["image"][0], image, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(out["depth"][0], depth, rtol=1e-4, atol=1e-4)  # SealD: depth is NOT normalised
    torch.testing.assert_close(out["weights_sum"], ws, rtol=1e-4, atol=1e-4)
    # and the edit is visible: the plain D-NeRF render of the same field differs
    net.seal_mapper = None
    with torch.no_grad():
        plain = net.render(ro[None], rd[None], time, perturb=False, force_all_rays=True)
    assert float((plain["image"] - out["image"]).abs().max()) > 1e-3


def test_seal_pretrainer_builds_point_sets_and_fits_the_table(cuda_dev):
    """f4 (SealDNeRF/utils.py:386-562 + SealNeRF/trainer.py:363-462): the three pre-training sets of a bbox edit — local points are
    exactly the lattice points the mapping moves (labels = teacher at the mapped point, colours through map_color), surrounding /
    global points are the ones it leaves alone — and a few epochs of the fused pre-training step pull the student's field towards
    the labels with every MLP untouched."""
    sys.path.insert(0, HERE)
    import ref_cases as rc
    from seald_nerf_b200.SealDNeRF.pretrain import SealPretrainer
    from seald_nerf_b200.trainer import FusedTrainer
    teacher = rc.ours_model(cuda_dev, seald=True)
    mapper = seal_mapper_from_dict(rc.seal_mapper_dict("bbox"))
    teacher.init_mapper(mapper=mapper)
    student = rc.ours_model(cuda_dev, seald=True)
    student.init_mapper(mapper=mapper)
    tr = FusedTrainer(student, num_rays=1024, max_samples=8192, use_graph=False, train_deform=False, init_loss_scale=1024.0)
    w_before = [w.detach().clone() for w in student.mlp_weights()]
    pt = SealPretrainer(tr, teacher)
    t_edit = 0.3
    pt.init_pretraining(t_edit, epochs=4, batch_size=4096, lr=0.02, local_point_step=0.02, surrounding_point_step=0.05, global_point_step=0.25)
    data = pt.pretraining_data
    assert set(data) == {"local", "surrounding", "global"}
    for k, d in data.items():
        n = d["points"].shape[0]
        assert n > 0 and d["dirs"].shape == (n, 3) and d["sigma"].shape == (n,) and d["color"].shape == (n, 3)
        assert d["steps"][0] == 0 and d["steps"][-1] == n and all(b - a <= 4096 for a, b in zip(d["steps"], d["steps"][1:]))
        assert bool(torch.isfinite(d["sigma"]).all()) and bool(torch.isfinite(d["color"]).all())
        moved = mapper.map_mask(d["points"])
        assert bool(moved.all()) if k == "local" else not bool(moved.any())
    # local labels: the teacher evaluated at the mapped points
    p, dirs = data["local"]["points"][:512], data["local"]["dirs"][:512]
    mp, md, _ = mapper.map_to_origin(p, torch.zeros_like(p) + torch.tensor([1.0, 0.0, 0.0], device=cuda_dev))
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        s_t, c_t, _ = teacher(mp, md, torch.tensor([[t_edit]], device=cuda_dev))
    torch.testing.assert_close(data["local"]["sigma"][:512], s_t.float().reshape(-1), rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(data["local"]["color"][:512], mapper.map_color(mp, md, c_t.float()), rtol=0, atol=2e-3)
    losses = [pt.pretrain_one_epoch()["local"] for _ in range(4)]
    torch.cuda.synchronize()
    assert all(np.isfinite(losses)) and losses[-1] < 0.8 * losses[0], losses
    for a, b in zip(w_before, student.mlp_weights()):
        assert torch.equal(a, b.detach()), "pre-training must not touch the MLPs (freeze_mlp)"
    assert 0 < int(tr.step_dev) <= 4 * pt.local_step  # one optimiser step per slice, minus the ones GradScaler skipped


def _oracle_dict(mapper, kind):
    """A constructed seal_utils mapper as the dict oracle/seal.py evaluates."""
    d = {"type": kind, "map_triangles": mapper.map_triangles.numpy(), "map_test_dir": None if mapper.map_test_dir is None else mapper.map_test_dir.numpy()}
    for k, v in mapper.map_data.items():
        d[k] = np.asarray(v, np.float32) if not isinstance(v, (str, bool)) else v
    return d


@pytest.mark.parametrize("mode", ["dry", "linear"])
def test_brush_mapper_built_from_config_runs_like_the_oracle(cuda_dev, mode):
    """f4 mapper construction -> a18 runtime: a brush mapper built from a GUI-style config (get_seal_mapper) through the CUDA mapping
    vs oracle/seal.py evaluating the same tensors (oracle pinned to the reference's seal_utils.py by test_oracle_seal.py)."""
    sys.path.insert(0, os.path.dirname(HERE))
    import bench
    from oracle import seal as S
    m = bench.seald_mappers()["brush_" + mode]
    od = _oracle_dict(m, "brush")
    rng = np.random.default_rng(7)
    pts = (np.array([0.13, 0.08, 0.05]) + (rng.random((20000, 3)) * 2 - 1) * [0.05, 0.4, 0.12]).astype(np.float32)
    dirs = rng.normal(size=(20000, 3)).astype(np.float32)
    p, d, mask = m.map_to_origin(torch.from_numpy(pts).to(cuda_dev), torch.from_numpy(dirs).to(cuda_dev))
    po, do_, mo = S.map_to_origin(od, pts, dirs)
    assert 0.05 < mo.mean() < 0.9
    assert np.array_equal(mask.cpu().numpy(), mo)
    np.testing.assert_allclose(p.cpu().numpy(), po, rtol=1e-5, atol=2e-5)
