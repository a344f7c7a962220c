"""Megatron-LM GPT ``ParallelTransformerLayer`` (reference ``module_inject/containers/megatron_gpt.py``)."""
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

class MegatronLayerPolicy(TransformerPolicy):
    _orig_layer_class = None  # Megatron is not importable here: matched by structure
    version = 0
    moe_type = "standard"
    megatron_v2 = True
    use_mup = False

    def __init__(self, client_module, inference=True):
        super().__init__(inference, megatron_v2=MegatronLayerPolicy.megatron_v2, use_mup=MegatronLayerPolicy.use_mup)
        self.client_module = client_module

    @staticmethod
    def matches(module):
        a = getattr(module, "attention", None) or getattr(module, "self_attention", None)
        return a is not None and hasattr(a, "query_key_value") and hasattr(a, "dense") and \
            hasattr(module, "input_layernorm") and hasattr(getattr(module, "mlp", None), "dense_h_to_4h")

    def _attn(self):
        return getattr(self.client_module, "attention", None) or self.client_module.self_attention

    def get_hidden_heads(self):
        a = self._attn()
        heads = getattr(a, "num_attention_heads", None) or getattr(a, "num_attention_heads_per_partition")
        return a.query_key_value.weight.shape[1], heads, getattr(self.client_module.input_layernorm, "eps", 1e-5), \
            self.client_module.mlp.dense_h_to_4h.weight.shape[0]

    def attention(self):
        a = self._attn()
        return a.query_key_value.weight, a.query_key_value.bias, a.dense.weight, a.dense.bias

    def mlp(self, moe_type="standard", expert=0):
        m = self.client_module.mlp
        if hasattr(m, "deepspeed_moe"):
            e = m.deepspeed_moe.experts.deepspeed_experts[expert]
            return e.dense_h_to_4h.weight, e.dense_h_to_4h.bias, e.dense_4h_to_h.weight, e.dense_4h_to_h.bias
        return m.dense_h_to_4h.weight, m.dense_h_to_4h.bias, m.dense_4h_to_h.weight, m.dense_4h_to_h.bias

    def layernorm(self):
        m = self.client_module
        return m.post_attention_layernorm.weight, m.post_attention_layernorm.bias, m.input_layernorm.weight, \
            m.input_layernorm.bias


class DS_MegatronGPTContainer(MegatronContainer, BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_megatron_gpt import DeepSpeedMegatronGPTInference
        return DeepSpeedMegatronGPTInference
