"""Llama-2 reference-implementation layer (``TransformerBlock`` of the Meta code base: wq/wk/wv/wo, w1/w2/w3) (reference ``module_inject/containers/llama2.py``)."""
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

class LLAMA2LayerPolicy(TransformerPolicy):
    _orig_layer_class = None  # the Meta (non-HF) llama package is matched by structure, see ``matches``

    def __init__(self, client_module, inference=True):
        super().__init__(inference, mlp_act_func_type=ActivationFuncType.GATED_SILU, norm_type=NormType.RMSNorm)
        self.client_module = client_module

    @staticmethod
    def matches(module):
        a, f = getattr(module, "attention", None), getattr(module, "feed_forward", None)
        return all(hasattr(a, n) for n in ("wq", "wk", "wv", "wo")) and all(hasattr(f, n) for n in ("w1", "w2", "w3"))

    def get_hidden_heads(self):
        m = self.client_module
        a = m.attention
        heads = getattr(a, "n_local_heads", None) or getattr(a, "n_heads")
        return a.wq.weight.shape[1], heads, getattr(m.ffn_norm, "eps", 1e-6), m.feed_forward.w1.weight.shape[0]

    def num_kv_heads(self):
        a = self.client_module.attention
        return getattr(a, "n_local_kv_heads", None) or getattr(a, "n_kv_heads", -1) or -1

    def rotary(self):
        a = self.client_module.attention
        return a.wq.weight.shape[0] // (getattr(a, "n_local_heads", None) or a.n_heads), False, 10000.0

    def attention(self):
        a = self.client_module.attention
        return cat_qkv(a.wq.weight, a.wk.weight, a.wv.weight), None, a.wo.weight, None

    def mlp(self):
        f = self.client_module.feed_forward
        return torch.cat([f.w1.weight, f.w3.weight], 0), None, f.w2.weight, None

    def layernorm(self):
        m = self.client_module
        return m.ffn_norm.weight, None, m.attention_norm.weight, None


class DS_LLAMA2Container(MetaTensorContainer, BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_llama2 import DeepSpeedLlama2Inference
        return DeepSpeedLlama2Inference
