"""FusedTrainer (graph-captured train step) against the drop-in autograd path and basic training sanity."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(dev, seed=0):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    model = bench.build_scene(dev, seed)
    return model


def _batch(dev, n=4096, seed=0):
    from seald_nerf_b200 import synthetic as syn
    pose = syn.orbit_poses(4, dev, seed=seed)[1]
    g = torch.Generator().manual_seed(seed)
    inds = torch.randint(0, 640000, (n,), generator=g).to(dev)
    o, d = syn.get_rays(pose, syn.intrinsics(), 800, 800, inds)
    t = 0.43
    rgb, a = syn.render_gt(o, d, t, 128)
    return o, d, t, rgb + (1 - a).unsqueeze(-1)


def test_trainer_gradients_match_dropin_autograd(cuda_dev):
    """One forward+backward of the fused trainer == run_cuda + MSE + autograd through the drop-in modules."""
    from seald_nerf_b200.trainer import FusedTrainer
    model = _scene(cuda_dev)
    model.encoder.embeddings.data.uniform_(-0.3, 0.3)
    o, d, t, gt = _batch(cuda_dev)
    tr = FusedTrainer(model, num_rays=4096, max_samples=4096 * 48, perturb=False, init_loss_scale=128.0, use_graph=False)
    tr.set_inputs(o, d, t, gt)
    tr._forward_backward()
    torch.cuda.synchronize()
    m_live = int(tr.counter[0])
    assert 10000 < m_live <= tr.M
    g_table = tr.grad_table.clone() / 128.0
    g_w = [g.clone() / 128.0 for g in tr.grad_views]
    loss_fused = float(tr.loss)

    # drop-in path on the same parameters
    model.zero_grad()
    model.mean_count = tr.M - 128  # same sample budget
    model.local_step = 0
    with torch.autocast("cuda", dtype=torch.float16):
        out = model.render(o[None], d[None], torch.tensor([[t]], device=cuda_dev), bg_color=1, perturb=False, force_all_rays=False,
                           dt_gamma=0, max_steps=1024)
        loss = torch.nn.functional.mse_loss(out["image"][0], gt)
    (loss * 128.0).backward()
    assert abs(float(loss) - loss_fused) <= 1e-3 * max(1e-3, abs(loss_fused)) + 1e-6
    ref_table = model.encoder.embeddings.grad / 128.0
    scale = float(ref_table.abs().max())
    assert float((g_table - ref_table).abs().max()) <= 2e-2 * scale
    for gm, w in zip(g_w, model.mlp_weights()):
        gr = w.grad / 128.0
        assert float((gm - gr).abs().max()) <= 2e-2 * float(gr.abs().max()) + 1e-9


@pytest.mark.parametrize("use_graph", [False, True])
def test_training_reduces_loss(cuda_dev, use_graph):
    from seald_nerf_b200.trainer import FusedTrainer
    model = _scene(cuda_dev, seed=1)
    o, d, t, gt = _batch(cuda_dev, seed=2)
    tr = FusedTrainer(model, num_rays=4096, max_samples=4096 * 48, lr=1e-2, lr_net=1e-3, use_graph=use_graph)
    losses = []
    for i in range(60):
        tr.train_step(o, d, t, gt)
        if i % 10 == 0 or i == 59:
            losses.append(float(tr.loss))
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.6 * losses[0], losses
    assert int(tr.step_dev) >= 55  # a few steps may be skipped while the loss scale settles
    # host-input API returns the same kind of number
    l = tr.train_step_host(o.cpu(), d.cpu(), t, gt.cpu())
    assert np.isfinite(l)
    assert 8 <= tr.launches_per_step <= 13  # (11 launches on one GPU after the fusions)


def test_fused_composite_loss_kernel_equals_three_kernels(cuda_dev):
    """seald_composite_train_loss_fused == composite forward + mse_loss_bg + composite backward (same arithmetic and order):
    composited outputs bit-identical, gradients equal to fp32 round-off with the identical zero pattern behind each ray's
    early stop; loss equal up to the order of the per-CTA atomic partial sums."""
    from seald_nerf_b200.trainer import FusedTrainer
    res = []
    for fuse in (False, True):
        model = _scene(cuda_dev)
        model.encoder.embeddings.data.uniform_(-0.3, 0.3)
        model.density_scale = 200.0  # opaque enough that most rays hit the T < 1e-4 early stop
        o, d, t, gt = _batch(cuda_dev)
        tr = FusedTrainer(model, num_rays=4096, max_samples=4096 * 48, perturb=False, init_loss_scale=128.0, use_graph=False, fuse_composite=fuse)
        tr.grad_sigma.fill_(7.0); tr.grad_rgb.fill_(7.0)  # the fused kernel must overwrite every live sample row itself
        bg = torch.rand(4096, 3, device=cuda_dev)
        tr.set_inputs(o, d, t, gt, bg)
        tr._forward_backward()
        torch.cuda.synchronize()
        m = int(tr.counter[0])
        # the march hands out sample ranges to CTAs in arrival order: re-pack the per-sample gradients in ray order
        rays = tr.rays.cpu().numpy()
        gs, gc = tr.grad_sigma.cpu(), tr.grad_rgb.cpu()
        gs = torch.cat([gs[o_:o_ + k] for _, o_, k in rays])
        gc = torch.cat([gc[o_:o_ + k] for _, o_, k in rays])
        assert gs.shape[0] == m
        res.append((tr.image.clone(), tr.weights_sum.clone(), tr.depth.clone(), tr.pred.clone(), gs, gc, float(tr.loss), tr.grad_table.clone()))
    for a, b in zip(res[0][:4], res[1][:4]):
        assert torch.equal(a, b)  # image, weights_sum, depth, pred: bit-identical
    for a, b in zip(res[0][4:6], res[1][4:6]):  # gradients: same zero pattern (early stops), values to fp32 round-off (FMA contraction)
        assert torch.equal(a == 0, b == 0)
        torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-9)
    assert abs(res[0][6] - res[1][6]) <= 5e-6 * abs(res[0][6]) + 1e-9  # (fp32 atomic sum over the rays: order-dependent round-off)
    assert float((res[0][4] == 0).float().mean()) > 0.2  # many samples lie behind an early stop
    scale = float(res[0][7].abs().max())
    assert float((res[0][7] - res[1][7]).abs().max()) <= 1e-4 * scale  # (atomic order in the scatter)


def test_sample_buffer_overflow_leaves_no_stale_gradients(cuda_dev):
    """max_samples below the live sample count: rays whose range does not fit are dropped (raymarching.cu:408, :523), and the rows of
    the one range that straddles the end of the buffer must carry the reference's zero fill — not the previous step's samples and
    gradients.  Two steps on different batches with an overflowing buffer: the table / weight gradients of the second step must equal
    those of the same step run from clean buffers, for the fused compositing kernel and the three-kernel path alike."""
    from seald_nerf_b200.trainer import FusedTrainer
    grads = {}
    for fuse in (True, False):
        for poison in (True, False):
            model = _scene(cuda_dev)
            model.encoder.embeddings.data.uniform_(-0.3, 0.3)
            o, d, t, gt = _batch(cuda_dev, n=2048, seed=0)
            tr = FusedTrainer(model, num_rays=2048, max_samples=8192, perturb=False, init_loss_scale=128.0, use_graph=False, fuse_composite=fuse)
            if poison:  # what an earlier overflowing step would have left behind
                tr.xyzs.uniform_(-0.5, 0.5); tr.dirs.fill_(0.577); tr.deltas.fill_(0.01)
                tr.grad_sigma.fill_(3.0); tr.grad_rgb.fill_(-2.0)
            tr.set_inputs(o, d, t, gt)
            tr._forward_backward()
            torch.cuda.synchronize()
            live = int(tr.counter[0])
            assert live > tr.M, (live, tr.M)  # the batch really overflows
            rays = tr.rays.cpu()
            kept = (rays[:, 2] > 0) & (rays[:, 1] + rays[:, 2] <= tr.M)
            assert 0 < int(kept.sum()) < int((rays[:, 2] > 0).sum())
            straddle = (rays[:, 2] > 0) & (rays[:, 1] < tr.M) & (rays[:, 1] + rays[:, 2] > tr.M)
            if bool(straddle.any()):
                off = int(rays[straddle][0, 1])
                for buf in (tr.xyzs, tr.dirs, tr.deltas, tr.grad_sigma, tr.grad_rgb):
                    assert float(buf[off:tr.M].abs().max()) == 0.0
            grads[(fuse, poison)] = (tr.grad_table.clone(), [g.clone() for g in tr.grad_views], kept)
    for fuse in (True, False):
        (ta, wa, ka), (tb, wb, kb) = grads[(fuse, True)], grads[(fuse, False)]
        # (which rays fit depends on the order in which CTAs reserve their ranges: compare only when the same rays were kept)
        if torch.equal(ka, kb):
            scale = float(tb.abs().max())
            assert float((ta - tb).abs().max()) <= 1e-4 * scale
            for a, b in zip(wa, wb):
                assert float((a - b).abs().max()) <= 1e-3 * float(b.abs().max()) + 1e-9
        assert bool(torch.isfinite(ta).all())


@pytest.mark.gpu
def test_data_parallel_modes_two_gpus():
    """World-size-2 run of scripts/dp_check.py (needs 2 GPUs): the summed per-rank gradient equals the single-GPU gradient
    (rtol 1e-3), and 12 optimiser steps of every exchange mode — fused peer-memory kernels with and without NVSwitch multicast,
    NCCL reduce-scatter/all-gather, NCCL all-reduce; graph-captured and eager — follow the single-GPU loss trajectory (2e-3),
    skip no step, and leave fp16 table == half(fp32 master) on every rank."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(root, "scripts", "dp_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "MISMATCH" not in p.stdout and p.stdout.count(" OK") >= 9


def test_overflowing_loss_scale_skips_the_step_and_halves_the_scale(cuda_dev):
    """GradScaler semantics through the WHOLE step with the overflow flag raised by the kernels that produce the gradients (table
    scatter + seald_mlp_wgrad_umma_flag; the one-launch optimiser has no check pass of its own on one GPU): an absurd loss scale makes
    the fp16 gradients overflow -> nothing moves, the scale halves, the step counter stays; once the scale is sane the step is taken."""
    from seald_nerf_b200.trainer import FusedTrainer
    o, d, t, gt = _batch(cuda_dev, n=2048, seed=5)
    model = _scene(cuda_dev, seed=2)
    tr = FusedTrainer(model, num_rays=2048, max_samples=2048 * 48, lr=1e-2, lr_net=1e-3, perturb=False, init_loss_scale=2.0 ** 40,
                      growth_interval=1000, defer_table_update=False)
    p0 = tr.params.clone()
    tr.train_step(o, d, t, gt)
    torch.cuda.synchronize()
    assert int(tr.step_dev) == 0 and float(tr.loss_scale) == 2.0 ** 39 and torch.equal(tr.params, p0)
    assert float(tr.grads.abs().max()) == 0.0  # cleared for the next step either way
    for _ in range(40):
        tr.train_step(o, d, t, gt)
    torch.cuda.synchronize()
    assert int(tr.step_dev) >= 1 and float(tr.loss_scale) < 2.0 ** 30 and not torch.equal(tr.params, p0)
    assert bool(torch.isfinite(tr.params).all())


def test_fused_optimizer_tail_equals_five_kernel_tail_and_follows_lambda_lr(cuda_dev, monkeypatch):
    """The one-launch MLP tail (csrc/optim_tail.cu: overflow check -> Adam -> fp16 copies + tcgen05 tiles -> GradScaler.update ->
    lr_scheduler.step) against the round-1 sequence of five kernels on the SAME gradients and optimiser state: parameters, moments,
    fp16 staging copies and both packed operand layouts, step counter, loss scale and growth tracker agree — for a normal step, a
    step whose MLP gradient holds an inf (skipped, scale halved) and a step that reaches the growth interval; with lr_decay_iters
    the learning rate follows torch's LambdaLR(0.1 ** min(iter / iters, 1)) stepped every iteration."""
    import seald_nerf_b200._lib as L
    from seald_nerf_b200.trainer import FusedTrainer
    o, d, t, gt = _batch(cuda_dev, n=2048, seed=4)
    model = _scene(cuda_dev, seed=2)
    tr = FusedTrainer(model, num_rays=2048, max_samples=2048 * 48, lr=1e-2, lr_net=1e-3, perturb=False, init_loss_scale=1024.0, growth_interval=3,
                      use_graph=False, defer_table_update=False)
    assert tr.fused_tail
    for _ in range(3):  # some history in the moments
        tr.train_step(o, d, t, gt)
    ntp, nw = tr.n_table_pad, tr.n_weights
    for case in ("normal", "overflow", "growth"):
        tr.set_inputs(o, d, t, gt)
        tr.grads.zero_()
        tr._forward_backward()
        if case == "overflow":
            tr.grads[ntp + 12345] = float("inf")
        if case == "growth":
            tr.growth_tracker.fill_(2)
        state = [x.clone() for x in tr._state()] + [tr.grads.clone()]
        res = {}
        for fused in (True, False):
            for dst, src in zip(list(tr._state()) + [tr.grads], state):
                dst.copy_(src)
            tr.hw.refresh(tr.weight_views)
            tr.fused_tail = fused
            if not fused or tr._wgrad_flags():  # (one GPU: the weight-gradient kernel raises the flag, the tail trusts it)
                L.call("seald_grad_finite_check", tr.grads.data_ptr() + 4 * ntp, nw, L.ptr(tr.found_inf), L.stream())
            tr._optimizer()
            torch.cuda.synchronize()
            res[fused] = (tr.params.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone(), tr.hw.flat.clone(), tr.hw.packed_deform.clone(),
                          tr.hw.packed_deform_T.clone(), tr.table16.clone(), int(tr.step_dev), float(tr.loss_scale), int(tr.growth_tracker),
                          float(tr.grads.abs().max()))
        tr.fused_tail = True
        a, b = res[False], res[True]
        assert a[7:] == b[7:], (case, a[7:], b[7:])                         # step counter, loss scale, growth tracker, cleared gradients
        if case == "overflow":
            assert b[7] == int(state[7]) and b[8] == 0.5 * float(state[3]) and torch.equal(b[0], state[0])   # skipped: nothing moved, scale halved
        if case == "growth":
            assert b[8] == 2.0 * float(state[3]) and b[9] == 0
        for x, y in zip(a[:3], b[:3]):                                      # fp32 parameters / moments (table and MLP regions)
            torch.testing.assert_close(x, y, rtol=5e-5, atol=1e-9)      # (FMA contraction differs between the two kernels)
        for x, y in zip(a[3:7], b[3:7]):                                    # fp16 copies: casts of (almost) the same fp32 values
            assert float((x.float() - y.float()).abs().max()) <= 1e-3 * float(x.float().abs().max()) + 1e-8
            assert torch.equal(x == 0, y == 0)                              # same zero padding in every layout
        # the tiles the tail scattered == a fresh cast + pack of the fp32 weights it produced (the loop left the five-kernel result in
        # the buffers: put the fused tail's parameters back first — the two differ in the last fp32 bit here and there)
        tr.params.copy_(b[0])
        tr.hw.refresh(tr.weight_views)
        assert torch.equal(tr.hw.flat, b[3]) and torch.equal(tr.hw.packed_deform, b[4]) and torch.equal(tr.hw.packed_deform_T, b[5])

    # ---- learning-rate schedule on the device vs torch's LambdaLR
    monkeypatch.setenv("SEALD_FUSED_TAIL", "1")
    model = _scene(cuda_dev, seed=2)
    iters = 20
    tr = FusedTrainer(model, num_rays=2048, max_samples=2048 * 48, lr=1e-2, lr_net=1e-3, perturb=False, init_loss_scale=1024.0, lr_decay_iters=iters)
    ref_opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1, device=cuda_dev))], lr=1e-2)
    sched = torch.optim.lr_scheduler.LambdaLR(ref_opt, lambda it: 0.1 ** min(it / iters, 1))
    for i in range(25):
        assert tr.current_lr[0] == pytest.approx(sched.get_last_lr()[0], rel=1e-6)
        tr.train_step(o, d, t, gt)
        ref_opt.step()
        sched.step()
    torch.cuda.synchronize()
    assert int(tr.sched_step) == 25 and tr.current_lr[0] == pytest.approx(1e-3, rel=1e-6)


def test_ema_matches_torch_ema_rule(cuda_dev):
    """ema_update(): decay = min(0.95, (1 + n) / (10 + n)); shadow -= (1 - decay) * (shadow - param) (torch_ema, once per epoch)."""
    from seald_nerf_b200.trainer import FusedTrainer
    model = _scene(cuda_dev, seed=3)
    o, d, t, gt = _batch(cuda_dev, n=1024, seed=1)
    tr = FusedTrainer(model, num_rays=1024, max_samples=1024 * 48, lr=1e-2, lr_net=1e-3, perturb=False, init_loss_scale=1024.0, ema_decay=0.95)
    shadow = tr.params.clone()
    for epoch in range(3):
        for _ in range(4):
            tr.train_step(o, d, t, gt)
        tr.ema_update()
        decay = min(0.95, (1 + epoch + 1) / (10 + epoch + 1))
        tmp = (shadow - tr.params) * (1.0 - decay)
        shadow = shadow - tmp
    torch.testing.assert_close(tr.ema_shadow, shadow, rtol=1e-4, atol=1e-8)  # (the kernel contracts s - omd * (s - p) into an FMA)
    raw = tr.params.clone()
    tr.ema_copy_to()
    assert torch.equal(tr.params, tr.ema_shadow) and torch.equal(tr.table16, model.encoder.embeddings.data.half())
    tr.ema_restore()
    assert torch.equal(tr.params, raw)


def test_pipelined_host_steps_equal_synchronous_ones(cuda_dev):
    """train_step_host_pipelined (H2D of the next batch under the running step, loss read one step late) runs exactly the steps
    train_step_host runs: same loss sequence (shifted by one call), same parameters afterwards."""
    from seald_nerf_b200.trainer import FusedTrainer
    import bench
    ro, rd, ts, gt = bench.make_batches(4, cuda_dev, 0)
    ro, rd, gt = ro[:, :1024].contiguous().cpu(), rd[:, :1024].contiguous().cpu(), gt[:, :1024].contiguous().cpu()
    res = []
    for pipelined in (False, True):
        model = bench.build_scene(cuda_dev, seed=0)
        tr = FusedTrainer(model, num_rays=1024, max_samples=1024 * 48, lr=1e-2, lr_net=1e-3, perturb=False, init_loss_scale=1024.0)
        losses = []
        for i in range(6):
            fn = tr.train_step_host_pipelined if pipelined else tr.train_step_host
            losses.append(fn(ro[i % 4], rd[i % 4], float(ts[i % 4]), gt[i % 4]))
        if pipelined:
            assert losses[0] is None
            losses = losses[1:] + [tr.drain_host_pipeline()]
        tr.flush()
        torch.cuda.synchronize()
        res.append((losses, tr.params.clone()))
    (la, pa), (lb, pb) = res
    assert all(abs(a - b) <= 2e-3 * abs(a) + 1e-7 for a, b in zip(la, lb)), (la, lb)   # (atomic-order noise between two runs)
    assert float((pa - pb).abs().mean()) <= 0.05 * float(pa.abs().mean())
