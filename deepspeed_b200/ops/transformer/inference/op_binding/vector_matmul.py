"""``VectorMatMulOp`` (reference ``ops/transformer/inference/op_binding/vector_matmul.py``): plain projection ``x @ W^T`` (attention output / MLP output)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp, gemm_linear


class VectorMatMulOp(BaseOp):

    def forward(self, input: torch.Tensor, weight: torch.Tensor, async_op: bool = False):
        return gemm_linear(input, weight)
