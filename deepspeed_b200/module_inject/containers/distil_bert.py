"""DistilBERT encoder block (post-LN) (reference ``module_inject/containers/distil_bert.py``)."""
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

class HFDistilBertLayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.distilbert.modeling_distilbert", "TransformerBlock")

    def __init__(self, client_module, inference=False, preln=False):
        super().__init__(inference, pre_attn_norm=preln)
        self.client_module = client_module
        self.cuda_graph_supported = True

    def causal(self):
        return False

    def get_hidden_heads(self):
        a = self.client_module.attention
        return a.q_lin.weight.shape[1], a.n_heads, self.client_module.sa_layer_norm.eps, self.client_module.ffn.lin1.weight.shape[0]

    def attention(self):
        a = self.client_module.attention
        (qw, qb), (kw, kb), (vw, vb) = _wb(a.q_lin), _wb(a.k_lin), _wb(a.v_lin)
        return cat_qkv(qw, kw, vw), cat_qkv(qb, kb, vb), a.out_lin.weight, a.out_lin.bias

    def mlp(self):
        f = self.client_module.ffn
        return f.lin1.weight, f.lin1.bias, f.lin2.weight, f.lin2.bias

    def layernorm(self):
        m = self.client_module
        return m.output_layer_norm.weight, m.output_layer_norm.bias, m.sa_layer_norm.weight, m.sa_layer_norm.bias


class DS_DistilBERTContainer(BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_bert import DeepSpeedBERTInference
        return DeepSpeedBERTInference
