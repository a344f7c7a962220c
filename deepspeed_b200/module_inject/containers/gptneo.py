"""GPT-Neo block (unscaled attention, alternating local windows) (reference ``module_inject/containers/gptneo.py``)."""
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

class HFGPTNEOLayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.gpt_neo.modeling_gpt_neo", "GPTNeoBlock")

    def __init__(self, client_module, inference=True):
        super().__init__(inference, scale_attention=False)
        self.client_module = client_module
        self.act_name = "gelu_new"

    def get_hidden_heads(self):
        a = self.client_module.attn.attention
        return a.embed_dim, a.num_heads, self.client_module.ln_1.eps, self.client_module.mlp.c_fc.weight.shape[0]

    def local_window(self):
        a = self.client_module.attn
        if getattr(a, "attention_type", "global") != "local":
            return 0
        cfg = getattr(a.attention, "config", None) or self.hf_model_config
        return int(getattr(cfg, "window_size", 256))

    def attention(self):
        a = self.client_module.attn.attention
        return cat_qkv(a.q_proj.weight, a.k_proj.weight, a.v_proj.weight), None, a.out_proj.weight, a.out_proj.bias

    def mlp(self):
        m = self.client_module.mlp
        return m.c_fc.weight, m.c_fc.bias, m.c_proj.weight, m.c_proj.bias

    def layernorm(self):
        m = self.client_module
        return m.ln_2.weight, m.ln_2.bias, m.ln_1.weight, m.ln_1.bias


class DS_GPTNEOContainer(MetaTensorContainer, BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_gpt import DeepSpeedGPTInference
        return DeepSpeedGPTInference
