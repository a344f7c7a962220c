"""``MLPGemmOp`` (reference ``ops/transformer/inference/op_binding/mlp_gemm.py``): ``norm(input + residual + bias)`` -> GEMM -> activation -> GEMM; returns (output, residual_add)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp, gemm_linear


class MLPGemmOp(BaseOp):

    def forward(self, input, residual, weight_interm, weight_out, input_bias=None, bias=None, gamma=None, beta=None):
        c = self.config
        x = input if input_bias is None else input + input_bias
        if c.norm_type in ("rms", "rmsnorm"):
            normed, res = T.rms_norm(x, gamma, c.epsilon, residual=residual)
        else:
            normed, res = T.layer_norm(x, gamma, beta, c.epsilon, residual=residual)
        from ..ds_transformer import _mlp_act
        h = gemm_linear(normed, weight_interm, bias)
        return gemm_linear(_mlp_act(h, c.mlp_act_func_type), weight_out), res
