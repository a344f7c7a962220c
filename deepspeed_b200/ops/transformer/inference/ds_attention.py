"""Stand-alone fused self-attention block of the v1 inference layer (reference ``ops/transformer/inference/ds_attention.py``):
input norm + packed QKV GEMM -> cached attention -> output projection, assembled from the op bindings."""
import math

import torch
from torch import nn

from deepspeed_b200 import comm as dist

from .op_binding import LinearOp, QKVGemmOp, SoftmaxContextOp, VectorMatMulOp, WorkspaceOp


class DeepSpeedSelfAttention(nn.Module):
    num_layers = 0

    def __init__(self, config, mp_group=None, q_scales=None, q_groups=1, merge_count=1):
        super().__init__()
        self.config = config
        self.config.layer_id = DeepSpeedSelfAttention.num_layers
        DeepSpeedSelfAttention.num_layers += 1
        c = config
        tp = c.mp_size
        self.heads = c.heads // tp
        self.kv = (c.num_kv if c.num_kv > 0 else c.heads) // tp
        d = c.hidden_size // c.heads
        dt = c.dtype if c.dtype in (torch.float16, torch.bfloat16, torch.float32) else torch.float16
        p = lambda *s: nn.Parameter(torch.empty(*s, dtype=dt), requires_grad=False)
        self.attn_qkvw, self.attn_qkvb = p((self.heads + 2 * self.kv) * d, c.hidden_size), p((self.heads + 2 * self.kv) * d)
        self.attn_ow, self.attn_ob = p(c.hidden_size, self.heads * d), p(c.hidden_size)
        self.mp_group = mp_group
        self.norm_factor = math.sqrt(math.sqrt(d)) ** -1 if c.scale_attention else 1.0  # applied to q and k each
        self.qkv_func, self.linear_func = QKVGemmOp(c), LinearOp(c)
        self.score_context_func, self.vector_matmul_func = SoftmaxContextOp(c), VectorMatMulOp(c)
        self.workspace = WorkspaceOp(c)

    def compute_attention(self, qkv_out, input_mask, layer_past, alibi, is_prompt, token_idx, position_ids):
        no_masking = input_mask is None and not self.config.triangular_masking
        return self.score_context_func(qkv_out, input_mask, self.heads, self.kv, self.norm_factor, no_masking,
                                       self.config.layer_id, DeepSpeedSelfAttention.num_layers, alibi, is_prompt, token_idx,
                                       position_ids)

    def forward(self, input, input_mask=None, head_mask=None, layer_past=None, get_present=False, encoder_hidden_states=None,
                encoder_attention_mask=None, output_attentions=False, norm_w=None, norm_b=None, alibi=None, **kwargs):
        if self.config.pre_layer_norm:
            qkv, normed = self.qkv_func(input, self.attn_qkvw, self.attn_qkvb, norm_w, norm_b)
        else:
            qkv, normed = self.linear_func(input, self.attn_qkvw, self.attn_qkvb), input
        is_prompt = layer_past is None and input.shape[1] > 1 or WorkspaceOp.seen(self.config.layer_id) == 0
        ctx, k, v = self.compute_attention(qkv, input_mask, layer_past, alibi, is_prompt, None, None)
        out = self.vector_matmul_func(ctx, self.attn_ow)
        if self.mp_group is not None and dist.get_world_size(self.mp_group) > 1:
            dist.inference_all_reduce(out, group=self.mp_group)
        return out, k, v, ctx, normed


class BloomSelfAttention(DeepSpeedSelfAttention):
    """BLOOM variant: ALiBi bias supplied by the caller (``alibi`` [batch*heads, 1, seq])."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.config.bigscience_bloom = True
