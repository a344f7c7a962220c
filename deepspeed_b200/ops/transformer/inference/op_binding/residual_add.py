"""``ResidualAddOp`` (reference ``ops/transformer/inference/op_binding/residual_add.py``): the post-MLP residual combination (sequential or parallel-attention form)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class ResidualAddOp(BaseOp):

    def forward(self, hidden_state, residual, add_bias, attention_output=None, residual_add=None, attention_bias=None,
                final_bias=None):
        c = self.config
        out = hidden_state + residual
        if add_bias and final_bias is not None:
            out = out + final_bias
        if not c.mlp_after_attn and attention_output is not None:  # GPT-J / NeoX: attention and MLP share the input
            out = out + attention_output
            if attention_bias is not None:
                out = out + attention_bias
        return out
