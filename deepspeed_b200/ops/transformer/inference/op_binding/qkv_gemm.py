"""``QKVGemmOp`` (reference ``ops/transformer/inference/op_binding/qkv_gemm.py``): ``norm(input)`` then the packed QKV projection; returns (qkv, normed input)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp, gemm_linear


class QKVGemmOp(BaseOp):

    def forward(self, input, weight, bias=None, gamma=None, beta=None):
        c = self.config
        if c.norm_type in ("rms", "rmsnorm"):
            normed = T.rms_norm(input, gamma, c.epsilon)
        else:
            normed = T.layer_norm(input, gamma, beta, c.epsilon)
        return gemm_linear(normed, weight, bias), normed
