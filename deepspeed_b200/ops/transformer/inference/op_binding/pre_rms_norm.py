"""``PreRMSNormOp`` (reference ``ops/transformer/inference/op_binding/pre_rms_norm.py``): ``residual += vals; return rmsnorm(residual), residual``."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class PreRMSNormOp(BaseOp):

    def forward(self, vals, residual, gamma, epsilon=None):
        return T.rms_norm(vals, gamma, epsilon if epsilon is not None else self.eps, residual=residual)
