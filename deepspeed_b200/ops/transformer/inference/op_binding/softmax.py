"""``SoftmaxOp`` (reference ``ops/transformer/inference/op_binding/softmax.py``): masked / causal / ALiBi softmax over attention scores."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class SoftmaxOp(BaseOp):

    def forward(self, attn_scores, attn_mask=None, alibi=None, triangular=False, recompute=False, local_attention=False,
                window_size=0, async_op=False, layer_scale=1.0, head_offset=0):
        return M.attn_softmax(attn_scores, mask=attn_mask, alibi=alibi, scale=layer_scale, causal=triangular,
                              window=window_size if local_attention else 0)
