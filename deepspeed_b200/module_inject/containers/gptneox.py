"""GPT-NeoX layer (per-head fused QKV, partial rotary, parallel residual) (reference ``module_inject/containers/gptneox.py``)."""
import torch

from deepspeed_b200.utils.types import ActivationFuncType, NormType

from ..policy import TransformerPolicy, cat_qkv, deinterleave_qkv, transpose  # noqa: F401
from .base import BaseTransformerContainer
from .features import MegatronContainer, MetaTensorContainer  # noqa: F401


def _cls(module_path, name):
    try:
        import importlib
        return getattr(importlib.import_module(module_path), name)
    except Exception:
        return None


def _wb(lin):
    return lin.weight, getattr(lin, "bias", None)

class GPTNEOXLayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.gpt_neox.modeling_gpt_neox", "GPTNeoXLayer")
    version = 0

    def __init__(self, client_module, inference=True, megatron_v2=True, split_qkv=False):
        super().__init__(inference, megatron_v2=megatron_v2, split_qkv=split_qkv)
        self.client_module = client_module

    def _cfg(self):
        return getattr(self.client_module.attention, "config", None) or self.hf_model_config

    def get_hidden_heads(self):
        c = self._cfg()
        return c.hidden_size, c.num_attention_heads, self.client_module.input_layernorm.eps, \
            self.client_module.mlp.dense_h_to_4h.weight.shape[0]

    def rotary(self):
        c = self._cfg()
        pct = getattr(c, "rotary_pct", None)
        if pct is None:
            pct = (getattr(c, "rope_parameters", None) or {}).get("partial_rotary_factor", 1.0)
        theta = (getattr(c, "rope_parameters", None) or {}).get("rope_theta", getattr(c, "rotary_emb_base", 10000.0))
        return int((c.hidden_size // c.num_attention_heads) * pct), True, float(theta)

    def mlp_after_attn(self):
        return not getattr(self.client_module, "use_parallel_residual", True)

    def parallel_mlp_own_norm(self):
        return True

    def attention(self):
        a = self.client_module.attention
        h = self._cfg().num_attention_heads
        return deinterleave_qkv(a.query_key_value.weight, h), deinterleave_qkv(a.query_key_value.bias, h), a.dense.weight, \
            a.dense.bias

    def mlp(self):
        m = self.client_module.mlp
        return m.dense_h_to_4h.weight, m.dense_h_to_4h.bias, m.dense_4h_to_h.weight, m.dense_4h_to_h.bias

    def layernorm(self):
        m = self.client_module
        return m.post_attention_layernorm.weight, m.post_attention_layernorm.bias, m.input_layernorm.weight, \
            m.input_layernorm.bias


class DS_GPTNEOXContainer(MetaTensorContainer, BaseTransformerContainer):  # the policy already de-interleaves the per-head QKV

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_gpt import DeepSpeedGPTInference
        return DeepSpeedGPTInference
