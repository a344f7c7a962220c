"""OPT decoder layer (ReLU, learned positions handled by the embedding) (reference ``module_inject/containers/opt.py``)."""
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

class HFOPTLayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.opt.modeling_opt", "OPTDecoderLayer")

    def __init__(self, client_module, inference=True, use_load_prefix=True):
        super().__init__(inference, linear_layer=True, pre_attn_norm=bool(getattr(client_module, "do_layer_norm_before", True)),
                         use_load_prefix=use_load_prefix)
        self.client_module = client_module
        self.mlp_act_func_type = ActivationFuncType.ReLU
        act = type(getattr(client_module, "activation_fn", None)).__name__.lower()
        self.act_name = "gelu" if "gelu" in act else "relu"

    def get_hidden_heads(self):
        a = self.client_module.self_attn
        return a.embed_dim, a.num_heads, self.client_module.self_attn_layer_norm.eps, self.client_module.fc1.weight.shape[0]

    def attention(self):
        a = self.client_module.self_attn
        return cat_qkv(a.q_proj.weight, a.k_proj.weight, a.v_proj.weight), cat_qkv(a.q_proj.bias, a.k_proj.bias, a.v_proj.bias), \
            a.out_proj.weight, a.out_proj.bias

    def mlp(self):
        m = self.client_module
        return m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias

    def layernorm(self):
        m = self.client_module
        return m.final_layer_norm.weight, m.final_layer_norm.bias, m.self_attn_layer_norm.weight, m.self_attn_layer_norm.bias


class DS_OPTContainer(MetaTensorContainer, BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_opt import DeepSpeedOPTInference
        return DeepSpeedOPTInference
