"""FusedRenderer (preallocated full-frame inference loop) against the drop-in NeRFRenderer.run_cuda / SealD teacher
run_cuda: same kernels and schedule, so images must agree to fp32 round-off (rtol/atol 1e-5)."""
import numpy as np
import pytest
import torch

from helpers import camera_rays, scene_bitfield, seal_mapper_from_dict

pytestmark = pytest.mark.gpu


def _model(cuda_dev, seald=False):
    from seald_nerf_b200.dnerf.network import NeRFNetwork
    from seald_nerf_b200.SealDNeRF.network import NeRFNetwork as SealNet
    torch.manual_seed(0)
    cls = SealNet if seald else NeRFNetwork
    net = cls(encoding="hashgrid", bound=1, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10).to(cuda_dev)
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    bits, _ = scene_bitfield()
    net.density_bitfield[:] = torch.from_numpy(bits).to(cuda_dev)[None]
    net.eval()
    return net


@pytest.mark.parametrize("tval", [0.3, 0.0])
def test_fused_renderer_matches_run_cuda(cuda_dev, tval):
    from seald_nerf_b200.renderer_fused import FusedRenderer
    net = _model(cuda_dev)
    ro, rd = camera_rays(5000, seed=3, center_crop=260)
    ro, rd = torch.from_numpy(ro).to(cuda_dev), torch.from_numpy(rd).to(cuda_dev)
    time = torch.tensor([[tval]], device=cuda_dev)
    with torch.no_grad():
        ref = net.render(ro[None], rd[None], time, perturb=False)
    fr = FusedRenderer(net, max_rays=6000)
    out = fr.render(ro[None], rd[None], time)
    assert out["image"].shape == (1, 5000, 3) and fr.iterations > 3 and fr.samples > 5000
    torch.testing.assert_close(out["image"], ref["image"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(out["depth"], ref["depth"], rtol=1e-5, atol=1e-5)
    # float time takes the host-side frame index
    out2 = fr.render(ro, rd, tval)
    torch.testing.assert_close(out2["image"], ref["image"][0], rtol=1e-5, atol=1e-5)
    # a second call reuses every buffer
    out3 = fr.render(ro[:1234], rd[:1234], time)
    torch.testing.assert_close(out3["image"], ref["image"][0, :1234], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kind", ["bbox", "brush", "anchor"])
def test_fused_renderer_seald_teacher(cuda_dev, kind):
    from oracle import seal as S
    from seald_nerf_b200.renderer_fused import FusedRenderer
    net = _model(cuda_dev, seald=True)
    if kind == "bbox":
        mp = S.make_bbox_mapper(center=(0.0, 0.1, 0.0), half=(0.2, 0.25, 0.2), translate=(0.08, 0.0, 0.03), rot_deg=30.0, hsv=(0.1, 0.0, 0.0))
    elif kind == "brush":
        mp = S.make_brush_mapper(mode="linear", pressure=0.05, depth=0.6, attenuation=0.03, rgb=(1.0, 0.0, 0.0))
    else:
        mp = S.make_anchor_mapper()
    net.init_mapper(mapper=seal_mapper_from_dict(mp))
    ro, rd = camera_rays(3000, seed=4, center_crop=200)
    ro, rd = torch.from_numpy(ro).to(cuda_dev), torch.from_numpy(rd).to(cuda_dev)
    time = torch.tensor([[0.55]], device=cuda_dev)
    with torch.no_grad():
        ref = net.render(ro[None], rd[None], time, perturb=False, force_all_rays=True)
    out = FusedRenderer(net, max_rays=3000).render(ro[None], rd[None], time)
    torch.testing.assert_close(out["image"], ref["image"], rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(out["depth"], ref["depth"], rtol=1e-5, atol=2e-5)   # raw depth (SealD does not normalise)
    torch.testing.assert_close(out["weights_sum"], ref["weights_sum"], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("kind", ["plain", "bbox", "brush"])
def test_one_pass_small_batch_render_matches_round_loop(cuda_dev, kind):
    """render_one_pass (training march + one field pass + composite_rays_train forward) against the round loop of render():
    identical sample positions; the early-stop test sits after the sample instead of before the next one, so a ray may take one
    sample fewer whose weight is below T_thresh.  Tolerance: image / weights_sum atol 2 * T_thresh, raw depth atol 4 * T_thresh
    (ray parameters are <= 3.5)."""
    from oracle import seal as S
    from seald_nerf_b200.renderer_fused import FusedRenderer
    seald = kind != "plain"
    net = _model(cuda_dev, seald=seald)
    if kind == "bbox":
        net.init_mapper(mapper=seal_mapper_from_dict(S.make_bbox_mapper(center=(0.0, 0.1, 0.0), half=(0.2, 0.25, 0.2), translate=(0.08, 0.0, 0.03),
                                                                        rot_deg=30.0, hsv=(0.1, 0.0, 0.0))))
    elif kind == "brush":
        net.init_mapper(mapper=seal_mapper_from_dict(S.make_brush_mapper(mode="linear", pressure=0.05, depth=0.6, attenuation=0.03, rgb=(1.0, 0.0, 0.0))))
    ro, rd = camera_rays(4096, seed=5, center_crop=220)
    ro, rd = torch.from_numpy(ro).to(cuda_dev), torch.from_numpy(rd).to(cuda_dev)
    time = torch.tensor([[0.41]], device=cuda_dev)
    for T_thresh in (1e-4, 1e-2):
        fr = FusedRenderer(net, max_rays=4096, min_samples=1 << 18)
        ref = fr.render(ro, rd, time, T_thresh=T_thresh)
        out = fr.render_one_pass(ro, rd, time, T_thresh=T_thresh)
        assert fr.iterations == 1 and fr.samples > 4096  # the one-pass path ran (no fallback) and marched real samples
        torch.testing.assert_close(out["image"], ref["image"], rtol=0, atol=2 * T_thresh)
        torch.testing.assert_close(out["weights_sum"], ref["weights_sum"], rtol=0, atol=2 * T_thresh)
        if seald:
            torch.testing.assert_close(out["depth"], ref["depth"], rtol=0, atol=4 * T_thresh)
        else:
            torch.testing.assert_close(out["depth"], ref["depth"], rtol=0, atol=4 * T_thresh)  # normalised by (far - near) ~ 3
        assert float((out["image"] - ref["image"]).abs().mean()) < 0.1 * T_thresh  # most rays agree to round-off
    # too few sample slots -> falls back to the round loop (same result as render())
    small = FusedRenderer(net, max_rays=4096, min_samples=4096)
    out_fb = small.render_one_pass(ro, rd, time, T_thresh=1e-4)
    ref_fb = small.render(ro, rd, time, T_thresh=1e-4)
    assert small.iterations > 1
    torch.testing.assert_close(out_fb["image"], ref_fb["image"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("kind", ["plain", "bbox"])
def test_sample_packed_rounds_equal_fixed_stride_rounds(cuda_dev, kind, monkeypatch):
    """The sample-packed round (samples of a round stored back to back, k_march_round_pack / k_composite_round_pack) against the
    reference's n_step-rows-per-ray layout: every ray composites the SAME sample sequence, so image / depth / weights are
    bit-identical whatever the first round's n_step — including n_step0 = 32 on a buffer of `rays` rows, where most CTAs of round 0 do
    not fit and their rays are deferred — while the field evaluates fewer rows."""
    from seald_nerf_b200.renderer_fused import FusedRenderer
    net = _model(cuda_dev, seald=(kind != "plain"))
    if kind != "plain":
        from oracle import seal as S
        net.init_mapper(mapper=seal_mapper_from_dict(S.make_bbox_mapper(center=(0.0, 0.1, 0.0), half=(0.2, 0.25, 0.2), translate=(0.08, 0.0, 0.03),
                                                                         rot_deg=30.0, hsv=(0.1, 0.0, 0.0))))
    ro, rd = camera_rays(70000, seed=5, center_crop=400)
    ro, rd = torch.from_numpy(ro).to(cuda_dev), torch.from_numpy(rd).to(cuda_dev)
    monkeypatch.setenv("SEALD_RENDER_PACK", "0")
    fr0 = FusedRenderer(net, max_rays=70000)
    ref = fr0.render(ro, rd, 0.4)
    assert not fr0.pack and fr0.samples > 70000
    monkeypatch.setenv("SEALD_RENDER_PACK", "1")
    for n0 in (1, 4, 32):
        monkeypatch.setenv("SEALD_RENDER_NSTEP0", str(n0))
        fr = FusedRenderer(net, max_rays=70000)
        assert fr.pack and fr.n_step0 == n0
        for rep in range(2):  # second call: the captured double-round graph is reused
            out = fr.render(ro, rd, 0.4)
            for k in ("image", "depth", "weights_sum"):
                assert torch.equal(out[k], ref[k]), (kind, n0, k, float((out[k] - ref[k]).abs().max()))
        assert fr.samples < fr0.samples, (fr.samples, fr0.samples)
    # perturbed start (noise per RAY, same generator state).  A perturbed eval render depends on how a ray's samples are cut into rounds
    # in the reference too (composite_rays advances rays_t by the deltas from the UNPERTURBED start, so every round restarts the
    # perturbation offset earlier, raymarching.cu:741,868): the two schedules agree to that offset's effect, not bit for bit
    monkeypatch.setenv("SEALD_RENDER_NSTEP0", "4")
    torch.manual_seed(77)
    ref_p = fr0.render(ro, rd, 0.4, perturb=True)
    fr = FusedRenderer(net, max_rays=70000)
    torch.manual_seed(77)
    out_p = fr.render(ro, rd, 0.4, perturb=True)
    assert not torch.equal(ref_p["image"], ref["image"])
    assert int((fr.noises[:70000] != 0).sum()) < 70000  # consumed (cleared) for every ray that was marched
    torch.testing.assert_close(out_p["image"], ref_p["image"], rtol=0, atol=2e-2)
    torch.testing.assert_close(out_p["weights_sum"], ref_p["weights_sum"], rtol=0, atol=4e-2)
    # opt-in coarse skipping (one bit per 8^3 block of cells): the same chain elements are sampled
    monkeypatch.setenv("SEALD_RENDER_COARSE", "1")
    monkeypatch.setenv("SEALD_RENDER_NSTEP0", "4")
    fr = FusedRenderer(net, max_rays=70000)
    assert fr.coarse is not None
    out = fr.render(ro, rd, 0.4)
    for k in ("image", "depth", "weights_sum"):
        assert torch.equal(out[k], ref[k]), (kind, "coarse", k)
    monkeypatch.setenv("SEALD_RENDER_COARSE", "0")
    # a row budget far below what round 0 asks for: most CTAs of the first rounds are deferred, the image does not change
    monkeypatch.setenv("SEALD_RENDER_NSTEP0", "32")
    monkeypatch.setenv("SEALD_RENDER_SLOTS_MULT", "0.25")
    fr = FusedRenderer(net, max_rays=70000, min_samples=1 << 12)
    out = fr.render(ro, rd, 0.4)
    assert fr.deferred > 0, "the budget was meant to overflow"
    for k in ("image", "depth", "weights_sum"):
        assert torch.equal(out[k], ref[k]), (kind, "deferred", k)
