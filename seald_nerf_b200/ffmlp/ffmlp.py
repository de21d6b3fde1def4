"""Drop-in `ffmlp` package: `FFMLP(input_dim, output_dim, hidden_dim, num_layers, activation='relu')` with the flat
fp32 `weights` Parameter, constraints, initialisation and padding rules of the reference (ffmlp/ffmlp.py:99-168), on the
fused tensor-core kernels of libseald_b200.so (fp16 operands, fp32 accumulation, no CUTLASS, no side streams).
"""
import math

import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from .. import _lib
from .._lib import ptr


def convert_activation(act):
    return {"relu": 0, "exponential": 1, "sine": 2, "sigmoid": 3, "squareplus": 4, "softplus": 5}.get(act, 6)


class _ffmlp_forward(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.half)
    def forward(ctx, inputs, weights, input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, inference=False,
                calc_grad_inputs=False):
        B = inputs.shape[0]
        inputs = inputs.contiguous()
        weights = weights.contiguous()
        if inputs.dtype != torch.half:
            inputs = inputs.half()
        if weights.dtype != torch.half:
            weights = weights.half()
        outputs = torch.empty(B, output_dim, device=inputs.device, dtype=inputs.dtype)
        if not inference:
            forward_buffer = torch.empty(num_layers, B, hidden_dim, device=inputs.device, dtype=inputs.dtype)
            _lib.call("seald_ffmlp_forward", ptr(inputs), ptr(weights), B, input_dim, output_dim, hidden_dim, num_layers, activation,
                      output_activation, ptr(forward_buffer), ptr(outputs), _lib.stream())
            ctx.save_for_backward(inputs, weights, forward_buffer)
            ctx.dims = (input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs)
        else:
            _lib.call("seald_ffmlp_forward", ptr(inputs), ptr(weights), B, input_dim, output_dim, hidden_dim, num_layers, activation,
                      output_activation, None, ptr(outputs), _lib.stream())
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        B = grad.shape[0]
        grad = grad.contiguous().half()
        inputs, weights, forward_buffer = ctx.saved_tensors
        input_dim, output_dim, hidden_dim, num_layers, activation, output_activation, calc_grad_inputs = ctx.dims
        grad_inputs = torch.empty_like(inputs) if calc_grad_inputs else None
        grad_weights = torch.zeros(weights.shape, device=weights.device, dtype=torch.float32)
        backward_buffer = torch.empty(num_layers, B, hidden_dim, device=grad.device, dtype=grad.dtype)
        _lib.call("seald_ffmlp_backward", ptr(grad), ptr(inputs), ptr(weights), ptr(forward_buffer), B, input_dim, output_dim, hidden_dim,
                  num_layers, activation, output_activation, ptr(backward_buffer), ptr(grad_inputs), ptr(grad_weights), _lib.stream())
        return grad_inputs, grad_weights, None, None, None, None, None, None, None, None  # autograd casts to the input dtype


ffmlp_forward = _ffmlp_forward.apply


class FFMLP(nn.Module):
    def __init__(self, input_dim, output_dim, hidden_dim, num_layers, activation="relu"):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        self.activation = convert_activation(activation)
        self.output_activation = convert_activation("none")
        self.tensorcore_width = 16

        assert hidden_dim in [16, 32, 64, 128, 256], f"FFMLP only support hidden_dim in [16, 32, 64, 128, 256], but got {hidden_dim}"
        assert input_dim > 0 and input_dim % 16 == 0, f"FFMLP input_dim should be 16 * m (m  > 0), but got {input_dim}"
        assert output_dim <= 16, f"FFMLP current only supports output dim <= 16, but got {output_dim}"
        assert num_layers >= 2, f"FFMLP num_layers should be larger than 2 (3 matmuls), but got {num_layers}"
        if hidden_dim == 256 or input_dim > 128 or self.activation != 0:
            raise NotImplementedError("seald_b200 FFMLP implements relu, hidden_dim <= 128 and input_dim <= 128")

        self.padded_output_dim = int(math.ceil(output_dim / 16)) * 16
        # one flat parameter buffer: [hidden x in][(num_layers - 1) x hidden x hidden][16 x hidden]  (ffmlp.py:121-122)
        self.num_parameters = hidden_dim * (input_dim + hidden_dim * (num_layers - 1) + self.padded_output_dim)
        self.weights = nn.Parameter(torch.zeros(self.num_parameters))
        self.reset_parameters()

    def cleanup(self):
        pass  # the reference frees split-K streams here; this implementation has no global state

    def __repr__(self):
        return (f"FFMLP: input_dim={self.input_dim} output_dim={self.output_dim} hidden_dim={self.hidden_dim} "
                f"num_layers={self.num_layers} activation={self.activation}")

    def reset_parameters(self):
        torch.manual_seed(42)  # the reference reseeds the global RNG here (ffmlp.py:142)
        std = math.sqrt(3 / self.hidden_dim)
        self.weights.data.uniform_(-std, std)

    def forward(self, inputs):
        B, C = inputs.shape
        # pad the batch to a multiple of 128; an aligned batch still gets one extra tile (ffmlp.py:157-159)
        pad = 128 - (B % 128)
        if pad > 0:
            inputs = torch.cat([inputs, torch.zeros(pad, C, dtype=inputs.dtype, device=inputs.device)], dim=0)
        outputs = ffmlp_forward(inputs, self.weights, self.input_dim, self.padded_output_dim, self.hidden_dim, self.num_layers, self.activation,
                                self.output_activation, not self.training, inputs.requires_grad)
        if B != outputs.shape[0] or self.padded_output_dim != self.output_dim:
            outputs = outputs[:B, :self.output_dim]
        return outputs
