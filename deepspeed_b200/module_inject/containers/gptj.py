"""GPT-J block (parallel attention + MLP, interleaved rotary) (reference ``module_inject/containers/gptj.py``)."""
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

class HFGPTJLayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.gptj.modeling_gptj", "GPTJBlock")

    def __init__(self, client_module, inference=True):
        super().__init__(inference, scale_attention=True)
        self.client_module = client_module
        self.act_name = "gelu_new"

    def get_hidden_heads(self):
        a = self.client_module.attn
        return a.embed_dim, a.num_attention_heads, self.client_module.ln_1.eps, self.client_module.mlp.fc_in.weight.shape[0]

    def rotary(self):
        a = self.client_module.attn
        return int(a.rotary_dim or a.head_dim), False, 10000.0

    def mlp_after_attn(self):
        return False

    def attention(self):
        a = self.client_module.attn
        return cat_qkv(a.q_proj.weight, a.k_proj.weight, a.v_proj.weight), None, a.out_proj.weight, None

    def mlp(self):
        m = self.client_module.mlp
        return m.fc_in.weight, m.fc_in.bias, m.fc_out.weight, m.fc_out.bias

    def layernorm(self):
        m = self.client_module
        return None, None, m.ln_1.weight, m.ln_1.bias


class DS_GPTJContainer(MetaTensorContainer, BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_gpt import DeepSpeedGPTInference
        return DeepSpeedGPTInference
