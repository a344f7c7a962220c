"""``MoEResMatmulOp`` (reference ``ops/transformer/inference/op_binding/moe_res_matmul.py``): residual-MoE mixing with
per-channel coefficients. ``coef`` arrives transposed, ``[..., 2 * hidden, 1]``: its first ``hidden`` entries weight the
residual (dense MLP) branch, the second ``hidden`` entries the expert output -- the layout the reference kernel reads
(``csrc/transformer/inference/csrc/gelu.cu: moe_res_matmul``)."""
import torch

from .base import BaseOp


class MoEResMatmulOp(BaseOp):

    def forward(self, residual: torch.Tensor, coef: torch.Tensor, output: torch.Tensor):
        c = coef.transpose(-1, -2)
        h = c.shape[-1] // 2
        return residual * c[..., :h] + output * c[..., h:]
