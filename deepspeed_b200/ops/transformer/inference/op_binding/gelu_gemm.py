"""``GELUGemmOp`` (reference ``ops/transformer/inference/op_binding/gelu_gemm.py``): ``gelu(x @ W1^T + b) @ W2^T`` (BERT-style MLP without the norm)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp, gemm_linear


class GELUGemmOp(BaseOp):

    def forward(self, input: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, weight_out: torch.Tensor):
        h = T.bias_act(gemm_linear(input, weight), bias, "gelu_tanh")
        return gemm_linear(h, weight_out)
