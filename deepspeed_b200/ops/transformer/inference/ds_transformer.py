"""Fused per-layer inference module with a contiguous KV cache.

API parity: reference ``model_implementations/transformers/ds_transformer.py:20 DeepSpeedTransformerInference``
(+ ``ops/transformer/inference/{ds_attention,ds_mlp}.py`` over the N8 kernels).  Kept for users that build
models out of individual fused layers; whole-model serving goes through the ragged engine.  The layer owns a
``[B_max, 2, kv_heads, max_out_tokens, d]`` cache, appends the new K/V each call and attends with flash SDPA
(prefill) or a single-query attention over the cache prefix (decode).
"""
import math

import torch
from torch import nn
import torch.nn.functional as F

from deepspeed_b200 import comm as dist
from deepspeed_b200.ops.kernels import transformer_ops as T
from .config import DeepSpeedInferenceConfig


def _act(x, name):
    name = str(name).lower()
    if "quick" in name:
        return x * torch.sigmoid(1.702 * x)
    if "relu" in name:
        return F.relu(x)
    if "silu" in name or "swish" in name:
        return F.silu(x)
    if "gelu" in name:
        return F.gelu(x, approximate="tanh" if "new" in name or "tanh" in name or "fast" in name else "none")
    return F.gelu(x)


class DeepSpeedTransformerInference(nn.Module):
    layer_id = 0

    def __init__(self, config: DeepSpeedInferenceConfig, mp_group=None, quantize_scales=None, quantize_groups=1,
                 merge_count=1, mlp_extra_grouping=False):
        super().__init__()
        self.config = config
        self.config.layer_id = DeepSpeedTransformerInference.layer_id
        DeepSpeedTransformerInference.layer_id += 1
        self.mp_group = mp_group
        c = config
        tp = c.mp_size
        h = c.hidden_size
        self.heads = c.heads // tp
        self.kv_heads = (c.num_kv if c.num_kv > 0 else c.heads) // tp
        self.d = h // c.heads
        dt = c.dtype if c.dtype in (torch.float16, torch.bfloat16, torch.float32) else torch.float16
        qkv_out = (self.heads + 2 * self.kv_heads) * self.d
        p = lambda *s: nn.Parameter(torch.empty(*s, dtype=dt), requires_grad=False)
        self.norm_w, self.norm_b = p(h), p(h)               # input norm
        self.attn_qkvw, self.attn_qkvb = p(qkv_out, h), p(qkv_out)
        self.attn_ow, self.attn_ob = p(h, self.heads * self.d), p(h)
        self.attn_nw, self.attn_nb = p(h), p(h)             # post-attention norm
        inter_rows = (2 if _is_gated(c.mlp_act_func_type) else 1) * (c.intermediate_size // tp)
        self.inter_w, self.inter_b = p(inter_rows, h), p(inter_rows)
        self.output_w, self.output_b = p(h, c.intermediate_size // tp), p(h)
        self.cache = None
        self.seen = 0
        self.rope = None
        if c.rotary_dim > 0:
            self.rope = T.RotaryTable(c.rotary_dim, c.max_out_tokens, base=c.rope_theta)

    def reset_cache(self):
        self.seen = 0

    def _norm(self, x, w, b, residual=None):
        if self.config.norm_type in ("rms", "rmsnorm"):
            return T.rms_norm(x, w, self.config.epsilon, residual=residual)
        return T.layer_norm(x, w, b, self.config.epsilon, residual=residual)

    def _reduce(self, x):
        if self.mp_group is not None and dist.get_world_size(self.mp_group) > 1:
            dist.inference_all_reduce(x, group=self.mp_group)
        return x

    def _attn(self, x, attn_mask):
        c = self.config
        B, S, _ = x.shape
        hq, hkv, d = self.heads, self.kv_heads, self.d
        qkv = F.linear(x, self.attn_qkvw, self.attn_qkvb).view(B, S, hq + 2 * hkv, d)
        q, k, v = qkv[:, :, :hq], qkv[:, :, hq:hq + hkv], qkv[:, :, hq + hkv:]
        if self.rope is not None:
            if self.rope.cos.device != x.device:
                self.rope.to(x.device)
            pos = torch.arange(self.seen, self.seen + S, device=x.device)
            q, k = (_rope if c.rotate_half or not c.rotate_every_two else _rope_interleaved)(q, k, self.rope, pos, c.rotary_dim)
        if self.cache is None or self.cache.shape[0] < B or self.cache.device != x.device:
            self.cache = torch.zeros(B, 2, hkv, c.max_out_tokens, d, dtype=x.dtype, device=x.device)
        if self.seen + S > c.max_out_tokens:
            raise RuntimeError(f"KV cache overflow: {self.seen + S} > max_out_tokens {c.max_out_tokens}")
        self.cache[:B, 0, :, self.seen:self.seen + S] = k.transpose(1, 2)
        self.cache[:B, 1, :, self.seen:self.seen + S] = v.transpose(1, 2)
        total = self.seen + S
        kk, vv = self.cache[:B, 0, :, :total], self.cache[:B, 1, :, :total]
        scale = 1.0 / math.sqrt(d) if c.scale_attention else 1.0
        window = c.window_size if c.local_attention else 0
        use_alibi = bool(c.bigscience_bloom)
        mask = attn_mask
        if mask is not None and mask.dim() == 2:  # [B, total] padding mask (1 = keep)
            mask = mask[:, None, None, :].bool() if mask.dtype != torch.bool else mask[:, None, None, :]
        if mask is not None and mask.shape[-1] != total:
            mask = mask[..., -total:] if mask.shape[-1] > total else None
        causal = c.triangular_masking and S > 1 and self.seen == 0 and mask is None and not window and not use_alibi
        need_struct = c.triangular_masking and (S > 1 or window or use_alibi) and not causal
        if need_struct or use_alibi:
            qi = torch.arange(S, device=x.device)[:, None] + self.seen
            kj = torch.arange(total, device=x.device)[None, :]
            keep = (kj <= qi) if c.triangular_masking else torch.ones(S, total, dtype=torch.bool, device=x.device)
            if window:
                keep = keep & (qi - kj < window)
            bias = torch.zeros(1, 1, S, total, dtype=torch.float32, device=x.device).masked_fill(~keep, float("-inf"))
            if use_alibi:
                if getattr(self, "_slopes", None) is None or self._slopes.device != x.device:
                    full = alibi_slopes(c.heads, x.device)
                    r = dist.get_rank(self.mp_group) if (self.mp_group is not None and c.mp_size > 1) else 0
                    self._slopes = full[r * hq:(r + 1) * hq]
                bias = bias + self._slopes.view(1, hq, 1, 1) * kj.to(torch.float32).view(1, 1, 1, total)
            if mask is not None:
                bias = bias.masked_fill(~mask, float("-inf")) if mask.dtype == torch.bool else bias + mask.float()
            mask = bias.to(q.dtype)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), kk, vv, attn_mask=mask, is_causal=causal, scale=scale,
                                           enable_gqa=hq != hkv)
        self.seen = total
        return o.transpose(1, 2).reshape(B, S, hq * d)

    def forward(self, input=None, input_mask=None, attention_mask=None, attn_mask=None, head_mask=None, layer_past=None,
                get_key_value=False, get_present=False, encoder_output=None, enc_dec_attn_mask=None, x=None,
                encoder_hidden_states=None, encoder_attention_mask=None, use_cache=False, alibi=None,
                output_attentions=False, layer_head_mask=None, past_key_value=None, **kwargs):
        c = self.config
        x = input if input is not None else x
        if x.dim() == 2:
            x = x.unsqueeze(0)
        mask = attention_mask if attention_mask is not None else (attn_mask if attn_mask is not None else input_mask)
        if not use_cache and layer_past is None and past_key_value is None and not get_present:
            self.seen = 0
        residual = x
        a_in = self._norm(x, self.norm_w, self.norm_b) if c.pre_layer_norm else x
        a = self._reduce(F.linear(self._attn(a_in, mask), self.attn_ow))
        if c.mlp_after_attn:
            if c.pre_layer_norm:
                f_in, residual = self._norm(a + self.attn_ob, self.attn_nw, self.attn_nb, residual=residual)
            else:
                residual = self._norm(residual + a + self.attn_ob, self.norm_w, self.norm_b)
                f_in = residual
            m = F.linear(_mlp_act(F.linear(f_in, self.inter_w, self.inter_b), c.mlp_act_func_type), self.output_w)
            out = residual + self._reduce(m) + self.output_b
            if not c.pre_layer_norm:
                out = self._norm(out, self.attn_nw, self.attn_nb)
        else:  # parallel attention + MLP (GPT-J / NeoX style)
            m_in = self._norm(x, self.attn_nw, self.attn_nb) if getattr(c, "parallel_mlp_own_norm", False) else a_in
            m = F.linear(_mlp_act(F.linear(m_in, self.inter_w, self.inter_b), c.mlp_act_func_type), self.output_w)
            out = residual + a + self.attn_ob + self._reduce(m) + self.output_b
        if c.return_single_tuple:
            return (out, )
        return (out, None) if c.return_tuple else out


def _is_gated(name):
    return "gated" in str(name).lower() or str(name) in ("3", "4", "ActivationFuncType.GATED_GELU", "ActivationFuncType.GATED_SILU")


def _mlp_act(h, name):
    """Activation of the first MLP GEMM's output; gated variants hold [gate; up] stacked on the feature dim."""
    if _is_gated(name):
        gate, up = h.chunk(2, dim=-1)
        return _act(gate, "silu" if "silu" in str(name).lower() or str(name).endswith("4") else "gelu") * up
    return _act(h, name)


def alibi_slopes(n_heads, device=None):
    """Per-head ALiBi slopes (geometric sequence; the closest power of two first, then the interleaved remainder)."""
    def pow2(n):
        start = 2.0**(-(2.0**-(math.log2(n) - 3)))
        return [start * (start**i) for i in range(n)]
    if math.log2(n_heads).is_integer():
        sl = pow2(n_heads)
    else:
        c = 2**math.floor(math.log2(n_heads))
        sl = pow2(c) + pow2(2 * c)[0::2][:n_heads - c]
    return torch.tensor(sl, dtype=torch.float32, device=device)


def _rope_interleaved(q, k, table, pos, rot_dim):
    """GPT-J style rotary: pairs are (0,1), (2,3), ... instead of (i, i + rot_dim/2)."""
    cos = table.cos[pos][None, :, None, :].to(torch.float32)
    sin = table.sin[pos][None, :, None, :].to(torch.float32)

    def rot(t):
        tf = t.float()
        a, b = tf[..., 0:rot_dim:2], tf[..., 1:rot_dim:2]
        r = torch.stack([a * cos - b * sin, b * cos + a * sin], dim=-1).flatten(-2)
        return torch.cat([r, tf[..., rot_dim:]], -1).to(t.dtype)

    return rot(q), rot(k)


def _rope(q, k, table, pos, rot_dim):
    half = rot_dim // 2
    cos = table.cos[pos][None, :, None, :].to(torch.float32)
    sin = table.sin[pos][None, :, None, :].to(torch.float32)

    def rot(t):
        tf = t.float()
        a, b = tf[..., :half], tf[..., half:rot_dim]
        out = torch.cat([a * cos - b * sin, b * cos + a * sin, tf[..., rot_dim:]], -1)
        return out.to(t.dtype)

    return rot(q), rot(k)
