"""Drop-in `activation.trunc_exp` (reference: activation.py:5-17).

Forward: exp evaluated in fp32 whatever the autocast dtype of the caller; backward: the incoming gradient times
exp(clamp(x, -15, 15)), i.e. the exponential is truncated only where it feeds the gradient.  Both directions are elementwise
kernels of libseald_b200.so (csrc/encoders.cu); inside the fused field kernels the same expression is part of the sigma head.
"""
import torch

from . import _lib
from ._lib import ptr


class TruncExp(torch.autograd.Function):
    """sigma = trunc_exp(h): CUDA tensors only (the library has no CPU path)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, h):
        _lib.require_cuda(h)
        h = h.contiguous()
        sigma = torch.empty_like(h)
        _lib.call("seald_trunc_exp_forward", ptr(h), ptr(sigma), h.numel(), _lib.stream())
        ctx.save_for_backward(h)
        return sigma

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_sigma):
        (h,) = ctx.saved_tensors
        grad_sigma = grad_sigma.float().contiguous()
        grad_h = torch.empty_like(h)
        _lib.call("seald_trunc_exp_backward", ptr(grad_sigma), ptr(h), ptr(grad_h), h.numel(), _lib.stream())
        return grad_h


def trunc_exp(h):
    return TruncExp.apply(h)
