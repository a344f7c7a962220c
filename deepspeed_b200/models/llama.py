"""Llama-family decoder (Llama-2/3, Mistral, Phi-3, Qwen2) built on the deepspeed_b200 kernels.

This is the flagship training model of the framework (``BASELINE.json``: Llama-3-8B ZeRO-3 bf16).
The reference ships no training model; its users bring HF ``LlamaForCausalLM`` whose hot path on
B200 is PyTorch ops.  B200-first choices here:

* packed ``qkv`` and ``gate_up`` projections -> 4 large GEMMs per layer instead of 7;
* residual-add fused into RMSNorm (one pass over the residual stream per norm);
* RoPE applied in place on the packed QKV buffer with precomputed fp32 tables (one launch for Q+K);
* SwiGLU as one kernel on the packed gate_up output;
* weight gradients are written **directly into the ZeRO flat gradient buffer** by the linear's
  backward (``out=`` GEMM), bypassing AccumulateGrad and the copy into the reduce bucket;
* LM head + cross entropy chunked over tokens with the softmax gradient produced in place, so the
  ``[tokens, vocab]`` logits never exist (role of the reference ``FPDT_LogitsLoss``,
  ``sequence/fpdt_layer.py:1137``);
* attention core through ``ops.attention`` (cuDNN/flash SDPA library path or the native sm_100a kernel).
"""
import math
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from deepspeed_b200.ops.kernels import transformer_ops as T
from deepspeed_b200.ops.attention import causal_attention
from deepspeed_b200.ops.linear import flat_linear


@dataclass
class LlamaConfig:
    vocab_size: int = 128256
    hidden_size: int = 4096
    intermediate_size: int = 14336
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 8
    head_dim: Optional[int] = None
    max_position_embeddings: int = 8192
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_scaling: Optional[dict] = None
    tie_word_embeddings: bool = False
    hidden_act: str = "silu"
    attention_bias: bool = False
    initializer_range: float = 0.02
    # training knobs
    checkpoint_layers: int = 0  # number of decoder layers (from the front) run with activation recompute
    loss_chunk_tokens: int = 2048
    attention_backend: str = "auto"

    def __post_init__(self):
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads

    @property
    def q_size(self):
        return self.num_attention_heads * self.head_dim

    @property
    def kv_size(self):
        return self.num_key_value_heads * self.head_dim

    def num_parameters(self, include_embeddings=True):
        h, i, L = self.hidden_size, self.intermediate_size, self.num_hidden_layers
        per = h * (self.q_size + 2 * self.kv_size) + self.q_size * h + 3 * h * i + 2 * h
        n = L * per + h
        if include_embeddings:
            n += self.vocab_size * h * (1 if self.tie_word_embeddings else 2)
        return n

    def flops_per_token(self, seq_len, causal=True):
        """Training FLOPs/token (fwd + bwd = 3x fwd matmul FLOPs; attention counted causal)."""
        h, i, L = self.hidden_size, self.intermediate_size, self.num_hidden_layers
        mm = 2 * (h * (self.q_size + 2 * self.kv_size) + self.q_size * h + 3 * h * i) * L + 2 * h * self.vocab_size
        attn = 4 * seq_len * self.q_size * L * (0.5 if causal else 1.0)
        return 3 * (mm + attn)


PRESETS = {
    "llama3-8b": dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=8, rope_theta=500000.0, max_position_embeddings=8192),
    "llama3-70b": dict(vocab_size=128256, hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                       num_attention_heads=64, num_key_value_heads=8, rope_theta=500000.0, max_position_embeddings=8192),
    "llama2-7b": dict(vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=32, rope_theta=10000.0, max_position_embeddings=4096),
    "phi3-mini": dict(vocab_size=32064, hidden_size=3072, intermediate_size=8192, num_hidden_layers=32,
                      num_attention_heads=32, num_key_value_heads=32, rope_theta=10000.0, max_position_embeddings=4096),
    "mistral-7b": dict(vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=8, rope_theta=10000.0, max_position_embeddings=32768),
    "tiny": dict(vocab_size=512, hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                 num_key_value_heads=2, max_position_embeddings=256, rope_theta=10000.0),
}


def llama_config(name: str, **overrides) -> LlamaConfig:
    d = dict(PRESETS[name])
    d.update(overrides)
    return LlamaConfig(**d)


class RMSNorm(nn.Module):

    def __init__(self, hidden, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden))
        self.eps = eps

    def forward(self, x, residual=None):
        return T.rms_norm(x, self.weight, self.eps, residual=residual)


class FlatLinear(nn.Module):
    """``nn.Linear`` whose backward writes dW straight into the ZeRO flat gradient buffer."""

    def __init__(self, in_features, out_features, bias=False):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None

    def forward(self, x):
        return flat_linear(x, self.weight, self.bias)


class LMHead(FlatLinear):
    """Output projection.  ``forward(h)`` -> logits; ``forward(h, labels)`` -> mean cross entropy computed
    chunk-wise without materialising the logits.  Being a module call (not a bare use of ``weight``) keeps
    the ZeRO-3 fetch hooks in the loop; its backward needs no weights (dW/dh are produced in forward)."""
    ds_skip_backward_fetch = True

    def forward(self, h, labels=None, chunk=2048, assumed_scale=1.0):
        if labels is None:
            return flat_linear(h, self.weight, self.bias)
        from deepspeed_b200.ops.linear import chunked_linear_xent
        return chunked_linear_xent(h.reshape(-1, h.shape[-1]), self.weight, labels.reshape(-1), chunk=chunk,
                                   assumed_scale=assumed_scale)


class LlamaAttention(nn.Module):

    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.cfg = cfg
        self.qkv_proj = FlatLinear(cfg.hidden_size, cfg.q_size + 2 * cfg.kv_size, bias=cfg.attention_bias)
        self.o_proj = FlatLinear(cfg.q_size, cfg.hidden_size, bias=False)

    def forward(self, x, rope, positions=None):
        cfg = self.cfg
        B, S, _ = x.shape
        qkv = self.qkv_proj(x)  # [B, S, (hq + 2 hkv) d]
        out = causal_attention(qkv.view(B * S, -1), B, S, cfg.num_attention_heads, cfg.num_key_value_heads,
                               cfg.head_dim, rope, positions, backend=cfg.attention_backend)
        return self.o_proj(out.view(B, S, cfg.q_size))


class LlamaMLP(nn.Module):

    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.gate_up_proj = FlatLinear(cfg.hidden_size, 2 * cfg.intermediate_size)
        self.down_proj = FlatLinear(cfg.intermediate_size, cfg.hidden_size)
        self.act = cfg.hidden_act

    def forward(self, x):
        if self.act == "silu" and self.gate_up_proj.bias is None and self.down_proj.bias is None:
            from deepspeed_b200.ops.linear import swiglu_mlp
            return swiglu_mlp(x, self.gate_up_proj.weight, self.down_proj.weight)
        return self.down_proj(T.gated_act(self.gate_up_proj(x), self.act))


class LlamaDecoderLayer(nn.Module):
    """Takes and returns ``(delta, residual)``: ``residual`` is the running stream *before* adding
    ``delta`` (the previous branch output), so every add is fused into the following RMSNorm."""

    def __init__(self, cfg: LlamaConfig, layer_idx: int):
        super().__init__()
        self.layer_idx = layer_idx
        self.input_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.self_attn = LlamaAttention(cfg)
        self.post_attention_layernorm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self.mlp = LlamaMLP(cfg)

    def forward(self, delta, residual, rope, positions=None):
        if residual is None:
            residual = delta
            h = self.input_layernorm(delta)
        else:
            h, residual = self.input_layernorm(delta, residual)
        attn = self.self_attn(h, rope, positions)
        h, residual = self.post_attention_layernorm(attn, residual)
        return self.mlp(h), residual


class LlamaModel(nn.Module):

    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.cfg = cfg
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.embed_tokens.ds_skip_backward_fetch = True  # embedding backward is an index-add: no weights needed
        self.layers = nn.ModuleList([LlamaDecoderLayer(cfg, i) for i in range(cfg.num_hidden_layers)])
        self.norm = RMSNorm(cfg.hidden_size, cfg.rms_norm_eps)
        self._rope = None

    def rope_table(self, device):
        if self._rope is None or self._rope.cos.device != torch.device(device):
            self._rope = T.RotaryTable(self.cfg.head_dim, self.cfg.max_position_embeddings, self.cfg.rope_theta, device,
                                       self.cfg.rope_scaling)
        return self._rope

    def forward(self, input_ids, positions=None):
        cfg = self.cfg
        h = self.embed_tokens(input_ids)
        rope = self.rope_table(h.device)
        delta, residual = h, None
        for i, layer in enumerate(self.layers):
            if self.training and i < cfg.checkpoint_layers and torch.is_grad_enabled():
                from deepspeed_b200.runtime.activation_checkpointing import checkpointing as ckpt
                delta, residual = ckpt.checkpoint(layer, delta, residual, rope, positions)
            else:
                delta, residual = layer(delta, residual, rope, positions)
        h, _ = self.norm(delta, residual)
        return h


class LlamaForCausalLM(nn.Module):
    """``forward(input_ids, labels=None)`` -> loss (if labels) or logits.  Labels are next-token
    targets aligned with ``input_ids`` (shifted inside, HF convention) unless ``shift_labels=False``."""

    def __init__(self, cfg: LlamaConfig):
        super().__init__()
        self.cfg = cfg
        self.model = LlamaModel(cfg)
        self.lm_head = LMHead(cfg.hidden_size, cfg.vocab_size)
        if cfg.tie_word_embeddings:
            self.lm_head.weight = self.model.embed_tokens.weight
        self.ds_loss_multiplier = 1.0  # set by the engine: loss_scale / grad_accum_steps
        self.apply(self._init_weights)

    def _init_weights(self, m):
        std = self.cfg.initializer_range
        if isinstance(m, FlatLinear):
            if m.weight.numel():
                nn.init.normal_(m.weight, mean=0.0, std=std)
        elif isinstance(m, nn.Embedding):
            if m.weight.numel():
                nn.init.normal_(m.weight, mean=0.0, std=std)

    def forward(self, input_ids, labels=None, positions=None, shift_labels=True):
        h = self.model(input_ids, positions)
        if labels is None:
            return self.lm_head(h)
        if shift_labels:
            labels = torch.cat([labels[:, 1:], torch.full_like(labels[:, :1], -100)], dim=1)
        return self.lm_head(h, labels=labels, chunk=self.cfg.loss_chunk_tokens, assumed_scale=self.ds_loss_multiplier)

    # ---- HF checkpoint interop ---------------------------------------------------------------------
    @staticmethod
    def convert_hf_state_dict(sd, cfg: LlamaConfig):
        """Pack HF ``q_proj/k_proj/v_proj`` and ``gate_proj/up_proj`` into this model's layout."""
        out = {}
        for k, v in sd.items():
            if ".self_attn.q_proj." in k:
                base = k.replace(".q_proj.", ".{}.")
                kk, vv = sd[base.format("k_proj")], sd[base.format("v_proj")]
                out[base.format("qkv_proj")] = torch.cat([v, kk, vv], dim=0)
            elif ".self_attn.k_proj." in k or ".self_attn.v_proj." in k:
                continue
            elif ".mlp.gate_proj." in k:
                up = sd[k.replace("gate_proj", "up_proj")]
                out[k.replace("gate_proj", "gate_up_proj")] = torch.cat([v, up], dim=0)
            elif ".mlp.up_proj." in k:
                continue
            elif "rotary_emb.inv_freq" in k:
                continue
            else:
                out[k] = v
        return out

    @torch.no_grad()
    def generate_greedy(self, input_ids, max_new_tokens=16):
        """Reference (no KV cache) greedy decode used by tests; the serving path is inference v2."""
        ids = input_ids
        for _ in range(max_new_tokens):
            logits = self(ids)[:, -1]
            ids = torch.cat([ids, logits.argmax(-1, keepdim=True)], dim=1)
        return ids
