"""``GELUGemmOp`` (reference ``ops/transformer/inference/op_binding/gelu_gemm.py``): ``gelu(x @ W1^T + b) @ W2^T`` (BERT-style MLP without the norm)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class GELUGemmOp(BaseOp):

    def forward(self, input: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, weight_out: torch.Tensor):
        h = T.bias_gelu(F.linear(input, weight), bias)
        return F.linear(h, weight_out)
