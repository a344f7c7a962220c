"""InternLM decoder layer (llama layout with attention biases) (reference ``module_inject/containers/internlm.py``)."""
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

from .llama import LLAMALayerPolicy


class InternLMLayerPolicy(LLAMALayerPolicy):
    _orig_layer_class = None  # remote-code model: matched by class name
    _also = []

    @staticmethod
    def matches(module):
        return type(module).__name__ in ("InternLMDecoderLayer", "InternLM2DecoderLayer")


class DS_InternLMContainer(MetaTensorContainer, BaseTransformerContainer):

    @property
    def layer_class(self):
        from deepspeed_b200.model_implementations.transformers.ds_llama2 import DeepSpeedLlama2Inference
        return DeepSpeedLlama2Inference
