"""``BiasAddOp`` (reference ``ops/transformer/inference/op_binding/bias_add.py``): ``activation + bias`` (in place on the kernel path)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class BiasAddOp(BaseOp):

    def forward(self, activation: torch.Tensor, bias: torch.Tensor):
        return T.bias_act(activation, bias, act=None)
