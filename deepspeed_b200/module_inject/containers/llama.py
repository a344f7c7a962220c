"""LLaMA-family decoder layer (RMSNorm, rotary, gated SiLU, GQA) (reference ``module_inject/containers/llama.py``)."""
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

class LLAMALayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.llama.modeling_llama", "LlamaDecoderLayer")
    _also = [_cls("transformers.models.mistral.modeling_mistral", "MistralDecoderLayer"),
             _cls("transformers.models.qwen2.modeling_qwen2", "Qwen2DecoderLayer")]

    def __init__(self, client_module, inference=True):
        super().__init__(inference, mlp_act_func_type=ActivationFuncType.GATED_SILU, norm_type=NormType.RMSNorm)
        self.client_module = client_module

    def _cfg(self):
        return getattr(self.client_module.self_attn, "config", None) or self.hf_model_config

    def get_hidden_heads(self):
        c, m = self._cfg(), self.client_module
        eps = getattr(m.input_layernorm, "variance_epsilon", getattr(m.input_layernorm, "eps", 1e-6))
        return c.hidden_size, c.num_attention_heads, eps, m.mlp.gate_proj.weight.shape[0]

    def num_kv_heads(self):
        c = self._cfg()
        return getattr(c, "num_key_value_heads", c.num_attention_heads)

    def rotary(self):
        c = self._cfg()
        theta = (getattr(c, "rope_parameters", None) or {}).get("rope_theta", getattr(c, "rope_theta", 10000.0))
        return getattr(c, "head_dim", None) or c.hidden_size // c.num_attention_heads, True, float(theta)

    def attention(self):
        a = self.client_module.self_attn
        (qw, qb), (kw, kb), (vw, vb) = _wb(a.q_proj), _wb(a.k_proj), _wb(a.v_proj)
        return cat_qkv(qw, kw, vw), (cat_qkv(qb, kb, vb) if qb is not None else None), a.o_proj.weight, a.o_proj.bias

    def mlp(self):
        m = self.client_module.mlp
        return torch.cat([m.gate_proj.weight, m.up_proj.weight], 0), None, m.down_proj.weight, None

    def layernorm(self):
        m = self.client_module
        return m.post_attention_layernorm.weight, None, m.input_layernorm.weight, None


class DS_LLAMAContainer(MetaTensorContainer, BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_llama2 import DeepSpeedLlama2Inference
        return DeepSpeedLlama2Inference
