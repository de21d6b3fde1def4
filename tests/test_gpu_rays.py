"""GPU parity of seald_get_rays_gather (csrc/train.cu) against oracle/rays.py and the golden outputs of the reference's own
get_rays (tests/golden/rays.npz).  Bar: rays_o and the gathered / copied ground truth bit-exact; rays_d rtol 1e-5 / atol 1e-6
(normalisation + 3x3 rotation in fp32 with a different FMA contraction than torch's kernels); blended ground truth 1e-6."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _call(d, poses, times, images, frame, inds, H, W, C, intr, bg=None):
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    N = inds.shape[0]
    ro, rd, gt = torch.empty(N, 3, device=d), torch.empty(N, 3, device=d), torch.empty(N, 3, device=d)
    t_out = torch.zeros(1, device=d)
    fr = torch.tensor([frame], dtype=torch.int32, device=d)
    _lib.call("seald_get_rays_gather", ptr(poses), ptr(times), ptr(images), ptr(fr), ptr(inds), N, H, W, C, float(intr[0]), float(intr[1]),
              float(intr[2]), float(intr[3]), ptr(bg), ptr(ro), ptr(rd), ptr(gt) if images is not None else None, ptr(t_out), _lib.stream())
    return ro, rd, gt, t_out


def test_get_rays_gather_vs_reference_golden_and_oracle(cuda_dev):
    from oracle import rays as orr
    d = cuda_dev
    g = np.load(os.path.join(GOLDEN, "rays.npz"))
    for k in range(2):
        c = {n[len("case%d_" % k):]: g[n] for n in g.files if n.startswith("case%d_" % k)}
        H, W, N = [int(v) for v in c["HWN"]]
        # the camera sits in slot 2 of a 3-frame table: exercises the frame indirection
        poses = torch.zeros(3, 4, 4, device=d)
        poses[2] = torch.from_numpy(c["pose"]).to(d)
        times = torch.tensor([0.1, 0.2, 0.7], device=d)
        inds = torch.from_numpy(c["inds"]).to(d)
        if c["image"].shape[0]:
            img = torch.zeros(3, H * W, 4, device=d)
            img[2] = torch.from_numpy(c["image"]).to(d)
        else:  # same generator call as make_rays_golden.py
            gen = torch.Generator().manual_seed(int(c["image_seed"][0]))
            img = torch.zeros(3, H * W, 4, device=d)
            img[2] = torch.rand(1, H, W, 4, generator=gen)[0].reshape(-1, 4).to(d)
        ro, rd, gt, t_out = _call(d, poses, times, img, 2, inds, H, W, 4, c["intr"])
        assert float(t_out) == float(np.float32(0.7))
        assert np.array_equal(ro.cpu().numpy(), c["rays_o"])
        np.testing.assert_allclose(rd.cpu().numpy(), c["rays_d"], rtol=1e-5, atol=1e-6)
        ro_o, rd_o = orr.get_rays(c["pose"], c["intr"], H, W, c["inds"])
        np.testing.assert_allclose(rd.cpu().numpy(), rd_o, rtol=1e-5, atol=1e-6)
        px = c["gt_rgba"]
        np.testing.assert_allclose(gt.cpu().numpy(), px[:, :3] * px[:, 3:] + (1 - px[:, 3:]), rtol=1e-6, atol=1e-6)
        # random background per ray, and RGB-only images are copied bit-exactly
        bg = torch.rand(N, 3, device=d)
        _, _, gt_bg, _ = _call(d, poses, times, img, 2, inds, H, W, 4, c["intr"], bg)
        np.testing.assert_allclose(gt_bg.cpu().numpy(), orr.gather_gt(img[2].cpu().numpy(), c["inds"], bg.cpu().numpy()), rtol=1e-6, atol=1e-6)
        img3 = img[:, :, :3].contiguous()
        _, _, gt3, _ = _call(d, poses, times, img3, 2, inds, H, W, 3, c["intr"])
        assert np.array_equal(gt3.cpu().numpy(), px[:, :3])


def test_trainer_step_from_resident_dataset(cuda_dev):
    """FusedTrainer.attach_dataset + train_step_frame: rays and targets generated inside the step graph from a frame index; the
    first step must equal a step on explicitly computed rays of the same pixels (same sampled indices)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from seald_nerf_b200 import synthetic as syn
    from seald_nerf_b200.trainer import FusedTrainer
    d = cuda_dev
    H = W = 200
    intr = syn.intrinsics(H, W)
    poses = syn.orbit_poses(4, d, seed=0)
    times = torch.tensor([0.0, 0.3, 0.6, 0.9], device=d)
    images = torch.rand(4, H * W, 3, device=d)
    res = []
    for mode in ("dataset", "explicit"):
        model = bench.build_scene(d)
        tr = FusedTrainer(model, num_rays=4096, max_samples=4096 * 32, perturb=False, init_loss_scale=128.0, use_graph=(mode == "dataset"))
        if mode == "dataset":
            tr.attach_dataset(poses, intr, H, W, images, times)
            torch.manual_seed(3)
            loss = float(tr.train_step_frame(2))
            inds = tr.inds.clone()
            assert int(inds.min()) >= 0 and int(inds.max()) < H * W and inds.unique().numel() > 3500  # 4096 draws from 40 000 pixels
            staged = (tr.rays_o.clone(), tr.rays_d.clone(), tr.gt.clone(), float(tr.time))
        else:
            ro, rdir = syn.get_rays(poses[2], intr, H, W, inds)
            loss = float(tr.train_step(ro, rdir, 0.6, images[2][inds]))
            # what the graph staged for itself == the explicitly computed batch of the same pixels
            assert torch.equal(staged[0], ro) and torch.equal(staged[2], images[2][inds]) and staged[3] == float(np.float32(0.6))
            torch.testing.assert_close(staged[1], rdir, rtol=1e-5, atol=1e-6)
        tr.flush()
        res.append((loss, tr.params[:tr.n_table].clone()))
    assert res[0][0] == pytest.approx(res[1][0], rel=1e-4)
    # one Adam step moves every touched entry by ~lr * sign(g): entries whose gradient is round-off noise around zero may differ by
    # 2 * lr, everything else must agree
    scale = float(res[1][1].abs().max())
    differ = ((res[0][1] - res[1][1]).abs() > 1e-3 * scale).float().mean()
    assert float(differ) < 1e-3
