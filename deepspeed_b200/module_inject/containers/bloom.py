"""BLOOM block (ALiBi, per-head fused QKV) (reference ``module_inject/containers/bloom.py``)."""
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

class BLOOMLayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.bloom.modeling_bloom", "BloomBlock")

    def __init__(self, client_module, inference=True, use_load_prefix=True, split_qkv=False):
        super().__init__(inference, linear_layer=True, use_load_prefix=use_load_prefix, split_qkv=split_qkv)
        self.client_module = client_module
        self.act_name = "gelu_tanh"

    def uses_alibi(self):
        return True

    def get_hidden_heads(self):
        a = self.client_module.self_attention
        return a.hidden_size, a.num_heads, self.client_module.input_layernorm.eps, \
            self.client_module.mlp.dense_h_to_4h.weight.shape[0]

    def attention(self):
        a = self.client_module.self_attention
        return deinterleave_qkv(a.query_key_value.weight, a.num_heads), deinterleave_qkv(a.query_key_value.bias, a.num_heads), \
            a.dense.weight, a.dense.bias

    def mlp(self):
        m = self.client_module.mlp
        return m.dense_h_to_4h.weight, m.dense_h_to_4h.bias, m.dense_4h_to_h.weight, m.dense_4h_to_h.bias

    def layernorm(self):
        m = self.client_module
        return m.post_attention_layernorm.weight, m.post_attention_layernorm.bias, m.input_layernorm.weight, \
            m.input_layernorm.bias


class DS_BloomContainer(MetaTensorContainer, BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_bloom import DeepSpeedBloomInference
        return DeepSpeedBloomInference
