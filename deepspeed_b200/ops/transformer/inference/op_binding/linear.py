"""``LinearOp`` (reference ``ops/transformer/inference/op_binding/linear.py``): ``x @ W^T + b`` (optionally producing the per-head layout the attention kernels read)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp, gemm_linear


class LinearOp(BaseOp):

    def forward(self, input, weight, bias=None, add_bias=True, do_flash_attn=False, num_heads=1, external_cache=None,
                num_layers=None):
        out = gemm_linear(input, weight, bias if add_bias else None)
        if do_flash_attn:
            b, s, _ = out.shape
            return out.view(b, s, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
        return out
