"""Memory-efficient linear for ZeRO-3 (reference ``runtime/zero/linear.py:41 LinearFunctionForZeroStage3``): the
backward re-reads the (possibly re-gathered) weight from the Parameter instead of keeping a gathered copy alive
in autograd's saved tensors."""
import math

import torch
from torch import nn
from torch.nn import init


class LinearFunctionForZeroStage3(torch.autograd.Function):

    @staticmethod
    def forward(ctx, input, weight, bias=None):
        ctx.save_for_backward(input, weight, bias)  # ``weight`` is the Parameter: its data is re-fetched in backward
        out = input.matmul(weight.t())
        if bias is not None:
            out = out + bias
        return out

    @staticmethod
    def backward(ctx, grad_output):
        input, weight, bias = ctx.saved_tensors
        gi = gw = gb = None
        if ctx.needs_input_grad[0]:
            gi = grad_output.matmul(weight)
        if ctx.needs_input_grad[1]:
            go2 = grad_output.reshape(-1, grad_output.shape[-1])
            gw = go2.t().matmul(input.reshape(-1, input.shape[-1]))
        if bias is not None and ctx.needs_input_grad[2]:
            gb = grad_output.reshape(-1, grad_output.shape[-1]).sum(0)
        return gi, gw, gb


def zero3_linear_wrap(input, weight, bias=None):
    return LinearFunctionForZeroStage3.apply(input, weight, bias)


class LinearModuleForZeroStage3(nn.Module):
    __constants__ = ["in_features", "out_features"]

    def __init__(self, in_features: int, out_features: int, bias: bool = True) -> None:
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features)) if bias else None
        if bias is False:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self) -> None:
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return LinearFunctionForZeroStage3.apply(input, self.weight, self.bias)

    def extra_repr(self) -> str:
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"


def print_rank_0(message, debug=False, force=False):
    from deepspeed_b200 import comm as dist
    if (debug or force) and (not dist.is_initialized() or dist.get_rank() == 0):
        print(message)
