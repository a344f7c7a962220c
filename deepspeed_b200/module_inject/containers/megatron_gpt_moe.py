"""Megatron-LM GPT layer whose MLP is a DeepSpeed MoE block (reference ``module_inject/containers/megatron_gpt_moe.py``)."""
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

from .base_moe import BaseTransformerMoEContainer
from .megatron_gpt import MegatronLayerPolicy


class MegatronMoELayerPolicy(MegatronLayerPolicy):

    def __init__(self, client_module, inference=True):
        super().__init__(client_module, inference)
        moe = getattr(client_module.mlp, "deepspeed_moe", None)
        self.num_experts = len(moe.experts.deepspeed_experts) if moe is not None else 1

    @staticmethod
    def matches(module):
        a = getattr(module, "attention", None) or getattr(module, "self_attention", None)
        return a is not None and hasattr(a, "query_key_value") and hasattr(getattr(module, "mlp", None), "deepspeed_moe")


class DS_MegatronGPTMoEContainer(MegatronContainer, BaseTransformerMoEContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_megatron_gpt import DeepSpeedMegatronGPTInference
        return DeepSpeedMegatronGPTInference
