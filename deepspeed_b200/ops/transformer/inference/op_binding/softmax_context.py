"""``SoftmaxContextOp`` (reference ``ops/transformer/inference/op_binding/softmax_context.py``): scores -> masked softmax -> context with the layer's KV cache appended in place; returns (context, key, value)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp
from .workspace import WorkspaceOp


class SoftmaxContextOp(BaseOp):

    def forward(self, query_key_value, attn_mask, heads, num_kv, norm_factor, no_masking, layer_id, num_layers, alibi=None,
                is_prompt=True, token_idx=None, position_ids=None):
        c = self.config
        b, s, _ = query_key_value.shape
        kv = num_kv if num_kv and num_kv > 0 else heads
        d = query_key_value.shape[-1] // (heads + 2 * kv)
        q, k, v = query_key_value.view(b, s, heads + 2 * kv, d).split([heads, kv, kv], dim=2)
        cache = WorkspaceOp.kv_cache(layer_id, b, kv, c.max_out_tokens, d, query_key_value)
        seen = 0 if is_prompt else WorkspaceOp.seen(layer_id)
        cache[:b, 0, :, seen:seen + s] = k.transpose(1, 2)
        cache[:b, 1, :, seen:seen + s] = v.transpose(1, 2)
        WorkspaceOp.set_seen(layer_id, seen + s)
        kk, vv = cache[:b, 0, :, :seen + s], cache[:b, 1, :, :seen + s]
        mask = None
        if not no_masking and (s > 1 or alibi is not None):
            qi = torch.arange(s, device=q.device)[:, None] + seen
            keep = torch.arange(seen + s, device=q.device)[None, :] <= qi
            mask = torch.zeros(s, seen + s, dtype=q.dtype, device=q.device).masked_fill(~keep, float("-inf"))
        if attn_mask is not None:
            am = attn_mask[..., -(seen + s):]
            mask = am if mask is None else mask + am
        if alibi is not None:
            al = alibi.view(b, heads, 1, -1)[..., :seen + s]
            mask = al if mask is None else mask + al
        ctx = F.scaled_dot_product_attention(q.transpose(1, 2), kk, vv, attn_mask=mask, scale=norm_factor * norm_factor
                                             if norm_factor else None, enable_gqa=heads != kv)
        return ctx.transpose(1, 2).reshape(b, s, heads * d), kk, vv
