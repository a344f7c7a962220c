"""Injection policies: how to read the weights of a model family's transformer layer (reference ``module_inject/policy.py``).

A policy wraps ONE original layer instance (``client_module``) and exposes its parameters in the canonical layout the
fused layer consumes: every weight is ``[out_features, in_features]``, the attention input projection is packed
``[q heads | k heads | v heads]`` on the output dim.
"""
from abc import ABC, abstractmethod

import torch

from deepspeed_b200.utils.types import ActivationFuncType, NormType


class DSPolicy(ABC):
    _orig_layer_class = None
    cuda_graph_supported = False

    def __init__(self):
        self.cuda_graph_supported = False

    @abstractmethod
    def attention(self):
        """-> (qkvw, qkvb, dense_w, dense_b)"""
        raise NotImplementedError


class TransformerPolicy(DSPolicy):
    hf_model_config = None  # set by the replacement driver to the model-level config

    def __init__(self, inference=True, linear_layer=True, scale_attention=True, megatron_v2=False, use_mup=False,
                 mlp_act_func_type=ActivationFuncType.GELU, pre_attn_norm=True, use_load_prefix=False, split_qkv=True,
                 norm_type=NormType.LayerNorm):
        super().__init__()
        self.inference = inference
        self.linear_layer = linear_layer  # False: weights are stored [in, out] (GPT-2 Conv1D)
        self.scale_attention = scale_attention
        self.is_megatron_v2 = megatron_v2
        self.use_mup = use_mup
        self.mlp_act_func_type = mlp_act_func_type
        self.pre_attn_norm = pre_attn_norm
        self.use_load_prefix = use_load_prefix
        self.split_qkv = split_qkv
        self.norm_type = norm_type

    @abstractmethod
    def get_hidden_heads(self):
        """-> (hidden_size, num_heads, layer_norm_eps, intermediate_size)"""
        raise NotImplementedError

    @abstractmethod
    def mlp(self):
        """-> (w1, b1, w2, b2); gated families return w1 = [gate; up]"""
        raise NotImplementedError

    @abstractmethod
    def layernorm(self):
        """-> (post_attention_norm_w, post_attention_norm_b, input_norm_w, input_norm_b)"""
        raise NotImplementedError

    # optional facts with defaults -----------------------------------------------------------------------------------
    def num_kv_heads(self):
        return -1

    def rotary(self):
        """-> (rotary_dim, rotate_half, rope_theta); rotary_dim <= 0: no rotary embedding"""
        return -1, False, 10000.0

    def local_window(self):
        return 0

    def mlp_after_attn(self):
        return True

    def parallel_mlp_own_norm(self):
        return False

    def uses_alibi(self):
        return False

    def causal(self):
        return True


# ---- layout helpers ---------------------------------------------------------------------------------------------------
def transpose(data):
    """Conv1D-style ``[in, out]`` weight -> ``[out, in]`` (contiguous)."""
    return data.t().contiguous()


def cat_qkv(q, k, v):
    return None if q is None else torch.cat([q, k, v], dim=0)


def deinterleave_qkv(w, heads, kv_heads=None):
    """Fused QKV stored per head as ``[heads, 3, head_dim, ...]`` (GPT-NeoX, BLOOM, Megatron v2) -> ``[q | k | v]``."""
    if w is None:
        return None
    d = w.shape[0] // (3 * heads)
    x = w.reshape(heads, 3, d, *w.shape[1:])
    return torch.cat([x[:, i].reshape(heads * d, *w.shape[1:]) for i in range(3)], dim=0)


def _transpose(x, heads=1, mp_replace=None):
    """Megatron v2 checkpoint order ``[heads, 3, d]`` -> the ``[3, heads, d]`` order of older Megatron / HF (reference name)."""
    return deinterleave_qkv(x, heads)


def maybe_copy(module, sd, weight_quantizer, mp_replace, dst_name, src_name, qkv=False, megatron_v2=False, split_qkv=False,
               heads=1):
    """Copy ``sd[src_name]`` into ``module.<dst_name>`` if present (checkpoint-driven loading)."""
    if src_name not in sd:
        return
    src = sd[src_name]
    if megatron_v2 and qkv:
        src = deinterleave_qkv(src, heads)
    dst = getattr(module, dst_name)
    with torch.no_grad():
        dst.copy_(src.to(dst.dtype).reshape(dst.shape))


def maybe_copy_qkv(module, sd, weight_quantizer, mp_replace, dst_name, src_names, split_qkv=False):
    if all(n in sd for n in src_names):
        dst = getattr(module, dst_name)
        with torch.no_grad():
            dst.copy_(torch.cat([sd[n] for n in src_names], dim=0).to(dst.dtype).reshape(dst.shape))


def maybe_copy_geglu(module, sd, weight_quantizer, mp_replace, dst_name, src_names):
    """Gated MLP: stack ``gate`` then ``up`` into the fused first projection."""
    maybe_copy_qkv(module, sd, weight_quantizer, mp_replace, dst_name, src_names)


def maybe_get_lora(p):
    """LoRA factors attached to a parameter by PEFT-style wrappers, as ``[right, left, scaling]`` (or [])."""
    if hasattr(p, "lora_right_weight"):
        return [p.lora_right_weight, p.lora_left_weight, p.lora_scaling]
    return []


def pack_lora_weights(p):
    return maybe_get_lora(p)
