"""``LayerNormOp`` (reference ``ops/transformer/inference/op_binding/layer_norm.py``): LayerNorm, optionally fused with a residual (+bias) add."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class LayerNormOp(BaseOp):

    def forward(self, vals, gamma, beta, epsilon=None):
        return T.layer_norm(vals, gamma, beta, epsilon if epsilon is not None else self.eps)

    @staticmethod
    def layer_norm_residual(vals, bias, res, gamma, beta, epsilon):
        x = vals if bias is None else vals + bias
        out, _ = T.layer_norm(x, gamma, beta, epsilon, residual=res)
        return out

    @staticmethod
    def layer_norm_residual_store_pre_ln_res(vals, bias, res, gamma, beta, epsilon):
        x = vals if bias is None else vals + bias
        out, pre = T.layer_norm(x, gamma, beta, epsilon, residual=res)
        return out, pre
