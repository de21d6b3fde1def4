"""Drop-in `gridencoder` package: `GridEncoder` / `grid_encode` with the constructor, attributes and forward
signature of the reference's gridencoder/grid.py (:96-161), backed by libseald_b200.so.

Differences that stay behind the seam:
  * outputs are produced directly as [B, L*C] (the reference writes [L,B,C] and permutes: grid.py:47,57);
  * `dy_dx` is not materialised: input gradients are recomputed from the table in backward with fp32
    accumulation (the reference stores [B, L*D*C] halfs and accumulates in fp16: gridencoder.cu:344-369);
  * gradients of the table are accumulated with warp-aggregated / shared-memory-privatised atomics.
Parameter / buffer names (`embeddings`, `offsets`) and shapes are the reference's, so checkpoints load.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from .. import _lib
from .._lib import ptr, F16, F32

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}


def _dtype_id(t):
    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.float32:
        return F32
    raise RuntimeError("GridEncoder supports fp16 / fp32 embeddings, got %s" % t.dtype)


class _grid_encode(Function):
    @staticmethod
    @custom_fwd(device_type="cuda")
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0):
        # inputs [B, D] fp32 in [0,1]; embeddings [sO, C]; offsets [L+1] int32 -> [B, L*C]
        _lib.require_cuda(inputs, embeddings, offsets)
        inputs = inputs.contiguous()
        if inputs.dtype != torch.float32:
            inputs = inputs.float()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)

        # half-precision table under autocast when C is even (reference: grid.py:43-44)
        if torch.is_autocast_enabled() and C % 2 == 0:
            embeddings = embeddings.to(torch.half)
        embeddings = embeddings.contiguous()

        outputs = torch.empty(B, L * C, device=inputs.device, dtype=embeddings.dtype)
        _lib.call("seald_grid_encode_forward", ptr(inputs), ptr(embeddings), ptr(offsets), ptr(outputs), None, B, D, C, L, S, H,
                  int(gridtype), int(bool(align_corners)), int(interpolation), _dtype_id(embeddings), None, _lib.stream())

        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.dims = [B, D, C, L, S, H, gridtype, interpolation]
        ctx.align_corners = align_corners
        ctx.calc_grad_inputs = calc_grad_inputs
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, embeddings, offsets = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation = ctx.dims
        grad = grad.contiguous()
        if grad.dtype != embeddings.dtype:
            grad = grad.to(embeddings.dtype)
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.empty_like(inputs) if ctx.calc_grad_inputs else None
        dt = _dtype_id(embeddings)
        _lib.call("seald_grid_encode_backward", ptr(grad), ptr(inputs), ptr(embeddings), ptr(offsets), ptr(grad_embeddings), None,
                  ptr(grad_inputs), B, D, C, L, S, H, int(gridtype), int(bool(ctx.align_corners)), int(interpolation), dt, dt, None, _lib.stream())
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype="hash", align_corners=False, interpolation="linear"):
        super().__init__()
        # the finest resolution, if given, overrides per_level_scale (reference: grid.py:100-102)
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))

        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners

        if input_dim not in (2, 3, 4, 5):
            raise RuntimeError("GridEncoding: input_dim must be 2, 3, 4 or 5")
        if level_dim not in (1, 2, 4, 8):
            raise RuntimeError("GridEncoding: C must be 1, 2, 4, or 8.")

        # rows per level: dense until the table limit, rounded up to a multiple of 8 (reference: grid.py:118-129)
        offsets = []
        offset = 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        self.register_buffer("offsets", torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim

        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        # inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, inputs.requires_grad,
                              self.gridtype_id, self.align_corners, self.interp_id)
        return outputs.view(prefix_shape + [self.output_dim])

    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        # TV regulariser: not called by the dnerf / SealD trainers -> out of the hot-path scope
        raise NotImplementedError("grad_total_variation is outside the D-NeRF/SealD hot path and is not provided")
