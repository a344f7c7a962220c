"""``EinsumSecSmEcmOp`` (reference ``ops/transformer/inference/op_binding/einsum_sec_sm_ecm.py``): MoE dispatch contraction ``einsum('sec,sm->ecm')``."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class EinsumSecSmEcmOp(BaseOp):

    def forward(self, Q: torch.Tensor, W: torch.Tensor):
        s, e, c = Q.shape
        return torch.matmul(Q.reshape(s, e * c).t().to(W.dtype), W).reshape(e, c, W.shape[-1])
