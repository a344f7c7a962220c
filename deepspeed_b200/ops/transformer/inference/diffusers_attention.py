"""Fused (cross-)attention for diffusers UNet / VAE blocks (reference ``ops/transformer/inference/diffusers_attention.py``):
packed QKV GEMM for self-attention, separate q / kv projections for cross-attention, flash SDPA, output projection."""
import torch
import torch.nn.functional as F
from torch import nn


class DeepSpeedDiffusersAttention(nn.Module):
    layer_id = 0

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.config.layer_id = DeepSpeedDiffusersAttention.layer_id
        DeepSpeedDiffusersAttention.layer_id += 1
        h = config.hidden_size
        dt = config.dtype if config.dtype in (torch.float16, torch.bfloat16, torch.float32) else torch.float16
        p = lambda *s: nn.Parameter(torch.empty(*s, dtype=dt), requires_grad=False)
        self.attn_qkvw, self.attn_qkvb = p(3 * h, h), p(3 * h)  # self-attention: packed
        self.attn_qw, self.attn_kw, self.attn_vw = p(h, h), p(h, h), p(h, h)  # cross-attention (context dim set on copy)
        self.attn_qb = None
        self.attn_ow, self.attn_ob = p(h, h), p(h)
        self.heads = config.heads
        self.do_out_bias = True

    def forward(self, input, context=None, input_mask=None):
        b, s, h = input.shape
        if context is None:
            q, k, v = F.linear(input, self.attn_qkvw, self.attn_qkvb if self.attn_qkvb is not None and
                               self.attn_qkvb.numel() else None).chunk(3, dim=-1)
        else:
            q, k, v = F.linear(input, self.attn_qw), F.linear(context, self.attn_kw), F.linear(context, self.attn_vw)
        sp = lambda t: t.reshape(b, -1, self.heads, t.shape[-1] // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v), attn_mask=input_mask)
        out = F.linear(o.transpose(1, 2).reshape(b, s, -1), self.attn_ow)
        return out + self.attn_ob if self.do_out_bias else out
