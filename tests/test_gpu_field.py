"""GPU parity of the fused field kernels (deform MLP, sigma/colour heads, weight-gradient GEMMs) and of FFMLP against
the torch-CPU oracle (oracle/field.py).

Tolerances (fp16 operands, fp32 accumulation): forward outputs rtol 2e-2 / atol 2e-3 against the fp16-emulating
oracle; gradients are compared relative to their largest magnitude: |g - g_ref| <= 6e-2 * max|g_ref| (fp16 activations AND fp16 activation gradients).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close_rel(a, b, tol, what):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.abs(b).max() + 1e-30
    err = np.abs(a - b).max() / scale
    assert err <= tol, "%s: max error %.3g of max |ref| %.3g (allowed %.3g)" % (what, err, scale, tol)


def _make_net(dev, seed=0, bound=1.0):
    from seald_nerf_b200.dnerf.network import NeRFNetwork
    torch.manual_seed(seed)
    net = NeRFNetwork(encoding="hashgrid", bound=bound, cuda_ray=True, density_scale=1, min_near=0.2, density_thresh=10).to(dev)
    # non-degenerate parameters: larger table values and a deformation net that actually moves points
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    return net


def _oracle_inputs(net):
    import oracle.field as of  # noqa: F401
    dw = [l.weight.detach().cpu().float() for l in net.deform_net]
    sw = [l.weight.detach().cpu().float() for l in net.sigma_net]
    cw = [l.weight.detach().cpu().float() for l in net.color_net]
    table = net.encoder.embeddings.detach().cpu().float()
    return dw, sw, cw, table


def _device_scales(net, dev):
    from test_gpu_grid import _device_scales as ds
    S = float(np.log2(net.encoder.per_level_scale))
    return ds(dev, net.encoder.offsets.cpu().numpy(), 3, 16, S, 16, net.encoder.gridtype_id, False)[0], S


@pytest.mark.parametrize("tval", [0.37, 0.0])
def test_field_forward_and_backward(cuda_dev, tval):
    import oracle.field as of
    net = _make_net(cuda_dev)
    net.train()
    M = 1000  # not a multiple of 128
    g = torch.Generator().manual_seed(1)
    xyz = (torch.rand(M, 3, generator=g) * 1.6 - 0.8)
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, generator=g), dim=-1)
    xyz[-5:] = 0; dirs[-5:] = 0  # zero padding rows as produced by the march
    time = torch.tensor([[tval]])
    sc_dev, S = _device_scales(net, cuda_dev)

    with torch.autocast("cuda", dtype=torch.float16):
        sigma, rgb, deform = net(xyz.to(cuda_dev), dirs.to(cuda_dev), time.to(cuda_dev))
    assert sigma.dtype == torch.float32 and sigma.shape == (M,) and rgb.shape == (M, 3) and deform.shape == (M, 3)

    dw, sw, cw, table = _oracle_inputs(net)
    params = [w.clone().requires_grad_(True) for w in dw + sw + cw]
    tab = table.clone().requires_grad_(True)
    nd, ns = len(dw), len(sw)
    offsets = net.encoder.offsets.cpu().numpy()
    so, ro, do_ = of.dnerf_forward(xyz, dirs, tval, params[:nd], params[nd:nd + ns], params[nd + ns:], tab, offsets, S, 16, 1.0, 1.0, True, 1, sc_dev)
    np.testing.assert_allclose(deform.detach().cpu().numpy(), do_.detach().numpy(), rtol=2e-2, atol=2e-3)
    # second oracle pass evaluated at the device's own (fp16-rounded) deformation: a 1e-3 shift of x moves fine-level
    # samples into neighbouring cells, which is not an error of the encoder or of the heads
    for q in params + [tab]:
        q.grad = None
    so, ro, do_ = of.dnerf_forward(xyz, dirs, tval, params[:nd], params[nd:nd + ns], params[nd + ns:], tab, offsets, S, 16, 1.0, 1.0, True, 1, sc_dev,
                                   deform_values=deform.detach().cpu())
    np.testing.assert_allclose(rgb.detach().cpu().numpy(), ro.detach().numpy(), rtol=2e-2, atol=3e-3)
    np.testing.assert_allclose(sigma.detach().cpu().numpy(), so.detach().numpy(), rtol=3e-2, atol=3e-3)
    if tval == 0.0:
        assert float(deform.abs().max()) == 0.0

    # backward with random upstream gradients
    # upstream gradients carry a loss scale of 256 (as under GradScaler) so the fp16 activation gradients stay normal
    gs = torch.randn(M, generator=g) * 0.1 * 256
    gc = torch.randn(M, 3, generator=g) * 256
    (sigma * gs.to(cuda_dev)).sum().add((rgb * gc.to(cuda_dev)).sum()).backward()
    (so * gs).sum().add((ro * gc).sum()).backward()
    mine = [w.grad for w in net.mlp_weights()]
    names = ["deform%d" % i for i in range(nd)] + ["sigma%d" % i for i in range(ns)] + ["color%d" % i for i in range(len(cw))]
    for name, gm, p in zip(names, mine, params):
        if tval == 0.0 and name.startswith("deform"):
            assert gm is None or float(gm.abs().max()) == 0.0  # no gradient reaches the deformation net at t == 0
            continue
        assert gm is not None, name
        _close_rel(gm.cpu().numpy(), p.grad.numpy(), 6e-2, name)
    _close_rel(net.encoder.embeddings.grad.cpu().numpy(), tab.grad.numpy(), 4e-2, "grid table")


def test_density_matches_forward(cuda_dev):
    net = _make_net(cuda_dev, seed=3)
    net.eval()
    M = 777
    xyz = (torch.rand(M, 3, device=cuda_dev) * 1.6 - 0.8)
    dirs = torch.nn.functional.normalize(torch.randn(M, 3, device=cuda_dev), dim=-1)
    with torch.no_grad():
        for tval in (0.6, 0.0):
            time = torch.tensor([[tval]], device=cuda_dev)
            sigma, _, deform = net(xyz, dirs, time)
            out = net.density(xyz, time)
            assert torch.equal(out["sigma"], sigma)
            if tval == 0.0:
                assert float(deform.abs().max()) == 0 and float(out["deform"].abs().max()) > 0  # density() still reports it
            else:
                assert torch.equal(out["deform"], deform)


@pytest.mark.parametrize("in_dim,hidden,layers,out_dim,B", [(32, 64, 2, 16, 1000), (16, 16, 3, 1, 128), (64, 128, 2, 3, 4000), (128, 32, 4, 16, 513)])
def test_ffmlp_vs_torch_mlp(cuda_dev, in_dim, hidden, layers, out_dim, B):
    """FFMLP semantics = the torch MLP of testing/test_ffmlp.py:11-43 (num_layers + 1 matrices, ReLU, no bias)."""
    import oracle.field as of
    from seald_nerf_b200.ffmlp import FFMLP
    net = FFMLP(in_dim, out_dim, hidden, layers).to(cuda_dev)
    assert net.weights.numel() == hidden * (in_dim + hidden * (layers - 1) + 16)
    flat = net.weights.detach().cpu()
    ws, o = [], 0
    ws.append(flat[o:o + hidden * in_dim].view(hidden, in_dim)); o += hidden * in_dim
    for _ in range(layers - 1):
        ws.append(flat[o:o + hidden * hidden].view(hidden, hidden)); o += hidden * hidden
    ws.append(flat[o:o + 16 * hidden].view(16, hidden))
    ws = [w.clone().requires_grad_(True) for w in ws]
    x = torch.randn(B, in_dim)
    xg = x.clone().requires_grad_(True)
    y_o = of.mlp(xg, ws, half=True)[:, :out_dim]

    net.train()
    xd = x.to(cuda_dev).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        y = net(xd)
    assert y.shape == (B, out_dim) and y.dtype == torch.float16
    np.testing.assert_allclose(y.float().detach().cpu().numpy(), y_o.detach().numpy(), rtol=2e-2, atol=2e-2)
    gy = torch.randn(B, out_dim)
    (y.float() * gy.to(cuda_dev)).sum().backward()
    (y_o * gy).sum().backward()
    gflat = torch.cat([w.grad.reshape(-1) for w in ws])
    _close_rel(net.weights.grad.cpu().numpy(), gflat.numpy(), 3e-2, "ffmlp weights")
    _close_rel(xd.grad.cpu().numpy(), xg.grad.numpy(), 3e-2, "ffmlp inputs")
    # inference variant gives the same outputs
    net.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        y2 = net(x.to(cuda_dev))
    assert torch.equal(y2, y.detach())
