"""GPU parity of the optimiser-side kernels of the training step (csrc/train.cu) against torch itself.

The reference trains with torch.optim.Adam(betas=(0.9, 0.99), eps=1e-15) under torch.cuda.amp.GradScaler (main_dnerf.py:129,
nerf/utils.py:879-886): k_adam / k_loss_scale_update restate exactly that, so the checker is torch.optim.Adam and GradScaler's
documented update rule on the same numbers.  Tolerance: fp32, rtol 2e-6 + atol 1e-9 per step (same formula, different fused order; |p| ~ 1e-2 so atol 2e-8 = a few ulp).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _adam_call(p, g, m, v, step_dev, scale, found, p16, zero_grad, lr=1e-2, betas=(0.9, 0.99), eps=1e-15, cap=None):
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    if cap is None:
        _lib.call("seald_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, betas[0], betas[1], eps, 1, ptr(step_dev), ptr(scale),
                  ptr(found), ptr(p16), int(zero_grad), _lib.stream())
    else:
        _lib.call("seald_adam_step_ex", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, betas[0], betas[1], eps, 1, ptr(step_dev), ptr(scale),
                  ptr(found), ptr(p16), int(zero_grad), int(cap), _lib.stream())


@pytest.mark.parametrize("n,cap", [(100003, None), (1 << 20, None), (1 << 20, 296), (8, None)])
def test_adam_matches_torch_optim(cuda_dev, n, cap):
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    torch.manual_seed(0)
    d = cuda_dev
    p0 = torch.randn(n, device=d) * 1e-2
    p = p0.clone()
    m, v = torch.zeros(n, device=d), torch.zeros(n, device=d)
    p16 = p0.half() if n % 4 == 0 else None
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    step_dev = torch.zeros(1, dtype=torch.int32, device=d)
    scale = torch.full((1,), 1024.0, device=d)
    found = torch.zeros(1, dtype=torch.int32, device=d)
    tracker = torch.zeros(1, dtype=torch.int32, device=d)
    never = torch.zeros(n, dtype=torch.bool, device=d)
    never[n // 3: n // 3 + n // 4] = True  # rows no sample ever reaches: gradient and moments stay exactly zero
    for it in range(6):
        g_true = torch.randn(n, device=d) * (10.0 ** np.random.default_rng(it).uniform(-6, 0))
        g_true[torch.rand(n, device=d) < 0.3] = 0.0  # untouched this step (but with history after step 0)
        g_true[never] = 0.0
        g = (g_true * 1024.0).contiguous()  # scaled gradients, un-scaled inside the kernel
        _adam_call(p, g, m, v, step_dev, scale, found, p16, zero_grad=True, cap=cap)
        _lib.call("seald_loss_scale_update", ptr(scale), ptr(found), ptr(tracker), 2.0, 0.5, 1000, ptr(step_dev), _lib.stream())
        ref.grad = (g_true * 1024.0) / 1024.0
        opt.step()
        assert float(g.abs().max()) == 0.0  # zero_grad
        torch.testing.assert_close(p, ref.data, rtol=2e-6 * (it + 1), atol=2e-8)  # |p| ~ 1e-2: a few ulp
    st = opt.state[ref]
    torch.testing.assert_close(m, st["exp_avg"], rtol=1e-5, atol=1e-7 * float(st["exp_avg"].abs().max()))
    torch.testing.assert_close(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-7 * float(st["exp_avg_sq"].abs().max()))
    assert int(step_dev) == 6 and int(tracker) == 6
    assert torch.equal(p[never], p0[never]) and float(m[never].abs().max()) == 0.0  # the skipped rows are exactly what Adam leaves
    if p16 is not None:
        assert torch.equal(p16, p.half())  # fp16 working copy refreshed in the same pass


def test_adam_p16_untouched_rows_keep_their_value(cuda_dev):
    """Rows with zero gradient and zero moments are not written at all: their fp16 copy keeps whatever the caller put there."""
    d = cuda_dev
    n = 4096
    p = torch.randn(n, device=d)
    p16 = p.half()
    g = torch.zeros(n, device=d)
    g[:1024] = 1.0
    m, v = torch.zeros(n, device=d), torch.zeros(n, device=d)
    step_dev = torch.zeros(1, dtype=torch.int32, device=d)
    scale = torch.ones(1, device=d)
    found = torch.zeros(1, dtype=torch.int32, device=d)
    before = p.clone()
    _adam_call(p, g, m, v, step_dev, scale, found, p16, zero_grad=True)
    assert torch.equal(p[1024:], before[1024:]) and torch.equal(p16[1024:], before[1024:].half())
    assert float((p[:1024] - before[:1024]).abs().min()) > 0 and torch.equal(p16[:1024], p[:1024].half())


def test_overflow_skips_update_and_backs_off_like_gradscaler(cuda_dev):
    """GradScaler semantics (torch/amp/grad_scaler.py: step() skips when found_inf, update(): scale *= backoff and the growth
    tracker resets; otherwise tracker += 1 and scale *= growth every growth_interval unskipped steps)."""
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    d = cuda_dev
    n = 1 << 16
    p = torch.randn(n, device=d)
    p_before = p.clone()
    m, v = torch.zeros(n, device=d), torch.zeros(n, device=d)
    g = torch.randn(n, device=d)
    g[12345] = float("inf")
    step_dev = torch.zeros(1, dtype=torch.int32, device=d)
    scale = torch.full((1,), 65536.0, device=d)
    found = torch.zeros(1, dtype=torch.int32, device=d)
    tracker = torch.full((1,), 7, dtype=torch.int32, device=d)
    _lib.call("seald_grad_finite_check", ptr(g), n, ptr(found), _lib.stream())
    assert int(found) != 0 and float(found.view(torch.float32)) == 1.0  # non-zero as an int, 1.0 as a float (summed over ranks)
    _adam_call(p, g, m, v, step_dev, scale, found, None, zero_grad=True)
    stash = torch.zeros(4, dtype=torch.int32, device=d)
    _lib.call("seald_loss_scale_update_stash", ptr(scale), ptr(found), ptr(tracker), 2.0, 0.5, 3, ptr(step_dev), ptr(stash), _lib.stream())
    assert torch.equal(p, p_before) and float(m.abs().max()) == 0 and float(g.abs().max()) == 0
    assert float(scale) == 32768.0 and int(tracker) == 0 and int(step_dev) == 0 and int(found) == 0
    assert int(stash[0]) != 0 and int(stash[1]) == 0 and float(stash[2:3].view(torch.float32)) == 65536.0
    # three clean steps: the scale doubles after the third (growth_interval = 3)
    for k in range(3):
        g.normal_()
        _lib.call("seald_grad_finite_check", ptr(g), n, ptr(found), _lib.stream())
        _adam_call(p, g, m, v, step_dev, scale, found, None, zero_grad=True)
        _lib.call("seald_loss_scale_update", ptr(scale), ptr(found), ptr(tracker), 2.0, 0.5, 3, ptr(step_dev), _lib.stream())
        assert int(step_dev) == k + 1
    assert float(scale) == 65536.0 and int(tracker) == 0


def test_step_begin_selects_the_frame_and_draws_fresh_uniform_noise(cuda_dev):
    """seald_step_begin: frame selection like seald_select_frame + loss reset + counter-based uniform noise in [0, 1) that changes with
    every launch (graph replays included) and is reproducible from the counter."""
    from seald_nerf_b200 import _lib
    from seald_nerf_b200._lib import ptr
    T, fb, N = 8, 4096, 4096
    bits = torch.randint(0, 255, (T, fb), dtype=torch.uint8, device=cuda_dev)
    occ = torch.rand(T, 6, device=cuda_dev)
    out = torch.zeros(fb, dtype=torch.uint8, device=cuda_dev); occ_out = torch.zeros(6, device=cuda_dev)
    counter = torch.full((2,), 9, dtype=torch.int32, device=cuda_dev); loss = torch.full((1,), 3.0, device=cuda_dev)
    noises = torch.zeros(N, device=cuda_dev); ctr = torch.zeros(2, dtype=torch.int64, device=cuda_dev)
    time = torch.tensor([0.63], device=cuda_dev)

    def call():
        _lib.call("seald_step_begin", ptr(time), T, ptr(bits), fb, ptr(out), ptr(occ), ptr(occ_out), ptr(counter), ptr(loss), ptr(noises), N,
                  ptr(ctr), _lib.stream())
    call()
    torch.cuda.synchronize()
    t_idx = int(0.63 * T)
    assert torch.equal(out, bits[t_idx]) and torch.equal(occ_out, occ[t_idx]) and int(counter.sum()) == 0 and float(loss) == 0.0
    a = noises.clone()
    assert float(a.min()) >= 0.0 and float(a.max()) < 1.0 and abs(float(a.mean()) - 0.5) < 0.02 and abs(float(a.std()) - 0.2887) < 0.01
    assert ctr.tolist() == [1, 0]
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        call()
        with torch.cuda.graph(g):
            call()
    torch.cuda.current_stream().wait_stream(s)
    draws = [noises.clone()]
    for _ in range(3):
        g.replay(); torch.cuda.synchronize()
        draws.append(noises.clone())
    assert int(ctr[0]) == 5 and all(not torch.equal(x, y) for x, y in zip(draws, draws[1:])) and not torch.equal(draws[0], a)
    # lag-1 correlation between consecutive draws and between neighbouring rays ~ 0
    c1 = float(torch.corrcoef(torch.stack([draws[1], draws[2]]))[0, 1]); c2 = float(torch.corrcoef(torch.stack([a[:-1], a[1:]]))[0, 1])
    assert abs(c1) < 0.06 and abs(c2) < 0.06
    ctr.zero_(); call(); torch.cuda.synchronize()
    assert torch.equal(noises, a)  # same counter -> same numbers
