"""CLIP text/vision encoder layer (pre-LN, quick-GELU) (reference ``module_inject/containers/clip.py``)."""
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

class HFCLIPLayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.clip.modeling_clip", "CLIPEncoderLayer")

    def __init__(self, client_module, inference=False):
        super().__init__(inference, pre_attn_norm=True, scale_attention=True)
        self.client_module = client_module
        self.cuda_graph_supported = True
        self.act_name = "quick_gelu" if "quick" in type(client_module.mlp.activation_fn).__name__.lower() else "gelu"

    def causal(self):
        return False  # the caller passes the (causal, for the text tower) mask explicitly

    def get_hidden_heads(self):
        a = self.client_module.self_attn
        return a.embed_dim, a.num_heads, self.client_module.layer_norm1.eps, self.client_module.mlp.fc1.weight.shape[0]

    def attention(self):
        a = self.client_module.self_attn
        return cat_qkv(a.q_proj.weight, a.k_proj.weight, a.v_proj.weight), cat_qkv(a.q_proj.bias, a.k_proj.bias, a.v_proj.bias), \
            a.out_proj.weight, a.out_proj.bias

    def mlp(self):
        m = self.client_module.mlp
        return m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias

    def layernorm(self):
        m = self.client_module
        return m.layer_norm2.weight, m.layer_norm2.bias, m.layer_norm1.weight, m.layer_norm1.bias


class DS_CLIPContainer(BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_gpt import DeepSpeedGPTInference
        return DeepSpeedGPTInference
