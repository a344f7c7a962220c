"""``VectorAddOp`` (reference ``ops/transformer/inference/op_binding/vector_add.py``): ``a + gamma * b``."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class VectorAddOp(BaseOp):

    def forward(self, a, b, gamma):
        return a + gamma * b
