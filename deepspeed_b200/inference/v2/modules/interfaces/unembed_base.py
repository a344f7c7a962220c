"""``DSUnembedBase`` interface + ``DSUnembedRegistry`` (reference ``modules/interfaces/unembed_base.py``)."""
from typing import Any, Dict, Type

import torch

from ..configs.unembed_config import DSUnembedConfig
from ..ds_module import DSModuleBase
from ..module_registry import DSModuleRegistryBase


class DSUnembedBase(DSModuleBase):

    @staticmethod
    def config_class() -> Type[DSUnembedConfig]:
        return DSUnembedConfig

    def __init__(self, config: DSUnembedConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)

    def forward(self, hidden_states, vocab_embedding, ragged_metadata, bias=None, gamma=None, beta=None) -> torch.Tensor:
        """Logits of the last token of every sequence: [n_sequences, vocab]."""
        raise NotImplementedError


class DSUnembedRegistry(DSModuleRegistryBase):

    @classmethod
    def associated_class(cls) -> Type[DSModuleBase]:
        return DSUnembedBase
