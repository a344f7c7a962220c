"""GPT-2 block (pre-LN, Conv1D weights stored [in, out]) (reference ``module_inject/containers/gpt2.py``)."""
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

class HFGPT2LayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.gpt2.modeling_gpt2", "GPT2Block")

    def __init__(self, client_module, inference=True):
        super().__init__(inference, linear_layer=False)  # HF GPT-2 uses Conv1D
        self.client_module = client_module
        self.act_name = "gelu_new"

    def get_hidden_heads(self):
        a = self.client_module.attn
        return a.embed_dim, a.num_heads, self.client_module.ln_1.eps, self.client_module.mlp.c_fc.weight.shape[1]

    def attention(self):
        a = self.client_module.attn
        return transpose(a.c_attn.weight), a.c_attn.bias, transpose(a.c_proj.weight), a.c_proj.bias

    def mlp(self):
        m = self.client_module.mlp
        return transpose(m.c_fc.weight), m.c_fc.bias, transpose(m.c_proj.weight), m.c_proj.bias

    def layernorm(self):
        m = self.client_module
        return m.ln_2.weight, m.ln_2.bias, m.ln_1.weight, m.ln_1.bias


class DS_GPT2Container(BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_gpt import DeepSpeedGPTInference
        return DeepSpeedGPTInference
