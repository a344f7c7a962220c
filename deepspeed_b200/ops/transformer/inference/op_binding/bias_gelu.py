"""``BiasGeluOp`` (reference ``ops/transformer/inference/op_binding/bias_gelu.py``): ``gelu(activation + bias)`` (tanh approximation, like the reference kernel)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class BiasGeluOp(BaseOp):

    def forward(self, activation: torch.Tensor, bias: torch.Tensor):
        # the reference kernel (csrc/transformer/inference/csrc/gelu.cu) evaluates the tanh form in fp32
        return T.bias_act(activation, bias, "gelu_tanh")
