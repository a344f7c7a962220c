"""``BiasResidualOp`` (reference ``ops/transformer/inference/op_binding/bias_residual.py``): ``output + bias + residual``."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class BiasResidualOp(BaseOp):

    def forward(self, output: torch.Tensor, residual: torch.Tensor, bias: torch.Tensor):
        return T.bias_residual(output, bias, residual)
