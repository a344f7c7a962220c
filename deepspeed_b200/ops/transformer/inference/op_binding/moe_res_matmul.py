"""``MoEResMatmulOp`` (reference ``ops/transformer/inference/op_binding/moe_res_matmul.py``): residual-MoE mixing ``mlp * coef[..., 0] + moe * coef[..., 1]``."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class MoEResMatmulOp(BaseOp):

    def forward(self, residual: torch.Tensor, coef: torch.Tensor, output: torch.Tensor):
        return residual * coef[..., 0:1] + output * coef[..., 1:2]
