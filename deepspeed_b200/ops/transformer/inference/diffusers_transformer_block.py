"""Fused diffusers ``BasicTransformerBlock`` (reference ``ops/transformer/inference/diffusers_transformer_block.py``):
norm1 -> self-attn -> norm2 -> cross-attn -> norm3 -> GEGLU feed-forward, residual + bias folded into the norms."""
import torch
import torch.nn.functional as F
from torch import nn

from deepspeed_b200.ops.kernels import transformer_ops as T

from .diffusers_2d_transformer import Diffusers2DTransformerConfig  # noqa: F401
from .op_binding import GatedActivationOp


class DeepSpeedDiffusersTransformerBlock(nn.Module):

    def __init__(self, equivalent_module: nn.Module, config=None):
        super().__init__()
        m = equivalent_module
        self.config = config
        ff = m.ff.net
        self.ff1_w, self.ff1_b = nn.Parameter(ff[0].proj.weight.detach(), False), nn.Parameter(ff[0].proj.bias.detach(), False)
        self.ff2_w, self.ff2_b = nn.Parameter(ff[2].weight.detach(), False), nn.Parameter(ff[2].bias.detach(), False)
        self.norm1_g, self.norm1_b, self.norm1_eps = m.norm1.weight, m.norm1.bias, m.norm1.eps
        self.norm2_g, self.norm2_b, self.norm2_eps = m.norm2.weight, m.norm2.bias, m.norm2.eps
        self.norm3_g, self.norm3_b, self.norm3_eps = m.norm3.weight, m.norm3.bias, m.norm3.eps
        self.attn_1, self.attn_2 = m.attn1, m.attn2
        self.gated = GatedActivationOp()

    def forward(self, hidden_states, context=None, timestep=None, encoder_hidden_states=None, **kwargs):
        ctx = context if context is not None else encoder_hidden_states
        x = hidden_states
        h = T.layer_norm(x, self.norm1_g, self.norm1_b, self.norm1_eps)
        x = x + self.attn_1(h)
        h = T.layer_norm(x, self.norm2_g, self.norm2_b, self.norm2_eps)
        x = x + self.attn_2(h, ctx) if ctx is not None else x + self.attn_2(h)
        h = T.layer_norm(x, self.norm3_g, self.norm3_b, self.norm3_eps)
        g = self.gated(F.linear(h, self.ff1_w), self.ff1_b, 3)  # GEGLU
        return x + F.linear(g, self.ff2_w, self.ff2_b)
