"""Base classes of the pluggable module layer of the ragged engine (reference ``inference/v2/modules/ds_module.py``).

A *module* = one logical block of a transformer (attention, embedding, linear, MoE, norm, unembedding) with possibly
several implementations; models are assembled from whatever implementation the heuristics pick for the run
configuration.  ``RaggedTransformer`` (model_implementations) is the fused whole-model composition of the same kernels;
these modules are the building blocks for custom architectures.
"""
from abc import ABC, abstractstaticmethod
from typing import Any, Dict, Type

import torch

from deepspeed_b200.runtime.config_utils import DeepSpeedConfigModel


class DSModuleConfig(DeepSpeedConfigModel):
    max_tokens: int = 2048


class DSModuleBase(torch.nn.Module, ABC):
    """``config_class``: the pydantic config it consumes; ``name()``: registry key; ``supports_config``: can this
    implementation run that configuration?"""

    @abstractstaticmethod
    def config_class() -> Type[DSModuleConfig]:
        ...

    @abstractstaticmethod
    def name() -> str:
        ...

    @abstractstaticmethod
    def supports_config(config: DSModuleConfig) -> bool:
        ...

    def __init__(self, config: DSModuleConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__()
        self._config = config
        self._implementation_config = implementation_config or {}

    @property
    def config(self):
        return self._config
