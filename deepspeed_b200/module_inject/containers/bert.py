"""BERT / RoBERTa encoder layer (post-LN, bidirectional) (reference ``module_inject/containers/bert.py``)."""
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

class HFBertLayerPolicy(TransformerPolicy):
    _orig_layer_class = _cls("transformers.models.bert.modeling_bert", "BertLayer")
    _also = [_cls("transformers.models.roberta.modeling_roberta", "RobertaLayer")]

    def __init__(self, client_module, inference=False):
        super().__init__(inference, pre_attn_norm=False)
        self.client_module = client_module
        self.cuda_graph_supported = True
        act = getattr(getattr(client_module.intermediate, "intermediate_act_fn", None), "__class__", type(None)).__name__.lower()
        self.act_name = "relu" if "relu" in act else ("gelu_new" if "new" in act or "tanh" in act else "gelu")

    def causal(self):
        return False

    def get_hidden_heads(self):
        a = self.client_module.attention.self
        q = a.query.weight
        ln = self.client_module.attention.output.LayerNorm
        return q.shape[1], a.num_attention_heads, ln.eps, self.client_module.intermediate.dense.weight.shape[0]

    def attention(self):
        a = self.client_module.attention
        (qw, qb), (kw, kb), (vw, vb) = _wb(a.self.query), _wb(a.self.key), _wb(a.self.value)
        return cat_qkv(qw, kw, vw), cat_qkv(qb, kb, vb), a.output.dense.weight, a.output.dense.bias

    def mlp(self):
        m = self.client_module
        return m.intermediate.dense.weight, m.intermediate.dense.bias, m.output.dense.weight, m.output.dense.bias

    def layernorm(self):
        m = self.client_module
        return m.output.LayerNorm.weight, m.output.LayerNorm.bias, m.attention.output.LayerNorm.weight, \
            m.attention.output.LayerNorm.bias


class DS_BERTContainer(BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_bert import DeepSpeedBERTInference
        return DeepSpeedBERTInference
