"""GPU parity of the frequency / SH encoders and trunc_exp (fp32; tolerance rtol 1e-5, atol 2e-6 for SH,
atol 2e-3 for __sinf at the highest octaves vs libm sin; exact against the reference's own kernels)."""
import numpy as np
import pytest
import torch

from conftest import load_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D,deg", [(3, 10), (1, 6), (3, 4), (2, 1)])
def test_freq_forward_backward(cuda_dev, D, deg):
    from seald_nerf_b200.freqencoder import FreqEncoder
    from oracle import encoders as oe
    enc = FreqEncoder(input_dim=D, degree=deg)
    torch.manual_seed(0)
    x = (torch.rand(5001, D, device=cuda_dev) * 2 - 1).requires_grad_(True)
    y = enc(x)
    assert y.shape == (5001, D + 2 * D * deg)
    y_o = oe.freq_encode(x.detach().cpu().numpy(), deg)
    np.testing.assert_allclose(y.detach().cpu().numpy(), y_o, rtol=0, atol=2e-3 if deg > 6 else 2e-5)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    gi_o = oe.freq_backward(g.cpu().numpy(), y.detach().cpu().numpy(), D, deg)
    np.testing.assert_allclose(x.grad.cpu().numpy(), gi_o, rtol=1e-4, atol=1e-3)
    ref = load_ref("freqencoder")
    if ref is not None:
        out_r = torch.empty_like(y)
        ref.freq_encode_forward(x.detach(), x.shape[0], D, deg, y.shape[1], out_r)
        assert torch.equal(out_r, y.detach())
        gi_r = torch.zeros_like(x)
        ref.freq_encode_backward(g, out_r, x.shape[0], D, deg, y.shape[1], gi_r)
        torch.testing.assert_close(gi_r, x.grad, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("deg", [1, 2, 3, 4, 5, 6, 7, 8])
def test_sh_forward_backward(cuda_dev, deg):
    from seald_nerf_b200.shencoder import SHEncoder
    from oracle import encoders as oe
    enc = SHEncoder(degree=deg)
    torch.manual_seed(1)
    d = torch.nn.functional.normalize(torch.randn(4097, 3, device=cuda_dev), dim=-1)
    d[0] = 0  # zero padding rows of the march output
    d = d.requires_grad_(True)
    y = enc(d)
    y_o = oe.sh_encode(d.detach().cpu().numpy(), deg)
    lo = 1 if deg > 4 else 0  # (the scipy-based oracle of bands 4..7 is defined for unit vectors: skip the zero row there)
    np.testing.assert_allclose(y.detach().cpu().numpy()[lo:], y_o[lo:], rtol=1e-5, atol=2e-5 if deg > 4 else 2e-6)
    g = torch.randn_like(y)
    (y * g).sum().backward()
    J = oe.sh_jacobian_fd(d.detach().cpu().numpy().astype(np.float64), deg)
    gi_o = np.einsum("bc,bdc->bd", g.cpu().numpy().astype(np.float64), J)
    if deg <= 4:  # (finite differences of the oracle leave the unit sphere: only meaningful for the polynomial restatement)
        np.testing.assert_allclose(d.grad.cpu().numpy(), gi_o, rtol=1e-4, atol=1e-4)
    ref = load_ref("shencoder")
    if ref is not None:
        out_r = torch.empty_like(y)
        dy_r = torch.empty(d.shape[0], 3 * deg * deg, device=cuda_dev)
        ref.sh_encode_forward(d.detach(), out_r, d.shape[0], 3, deg, dy_r)
        # bands 0..3: the same expressions; bands 4..7: z-polynomial x (x, y)-polynomial vs the reference's expanded forms (fp32 round-off)
        torch.testing.assert_close(out_r, y.detach(), rtol=1e-6 if deg <= 4 else 1e-5, atol=1e-7 if deg <= 4 else 2e-5)  # (measured 5e-6 on band-7 values of magnitude ~3)
        gi_r = torch.zeros_like(d)
        ref.sh_encode_backward(g, d.detach(), d.shape[0], 3, deg, dy_r, gi_r)
        torch.testing.assert_close(gi_r, d.grad, rtol=1e-5 if deg <= 4 else 1e-4, atol=1e-5 if deg <= 4 else 2e-4)


def test_trunc_exp(cuda_dev):
    from seald_nerf_b200.activation import trunc_exp
    from oracle import encoders as oe
    x = torch.linspace(-20, 10, 1001, device=cuda_dev).requires_grad_(True)  # exp(10) still fits the fp16 gradient
    with torch.autocast("cuda", dtype=torch.float16):
        y = trunc_exp(x.half())
    assert y.dtype == torch.float32
    y.sum().backward()
    xq = x.detach().half().float().cpu().numpy()
    np.testing.assert_allclose(y.detach().cpu().numpy(), oe.trunc_exp_forward(xq), rtol=1e-5)
    np.testing.assert_allclose(x.grad.cpu().numpy(), oe.trunc_exp_backward(xq, np.ones_like(xq)), rtol=2e-3, atol=1e-7)
