"""``DSEmbeddingBase`` interface + ``DSEmbeddingRegistry`` (reference ``modules/interfaces/embedding_base.py``)."""
from typing import Any, Dict, Type

import torch

from ..configs.embedding_config import DSEmbeddingsConfig
from ..ds_module import DSModuleBase
from ..module_registry import DSModuleRegistryBase


class DSEmbeddingBase(DSModuleBase):

    @staticmethod
    def config_class() -> Type[DSEmbeddingsConfig]:
        return DSEmbeddingsConfig

    def __init__(self, config: DSEmbeddingsConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)

    def transform_param(self, embed_param: torch.Tensor):
        return embed_param

    @property
    def output(self) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, ragged_batch, word_embeddings, position_embeddings=None, token_type_ids=None, token_type_embeddings=None):
        raise NotImplementedError


class DSEmbeddingRegistry(DSModuleRegistryBase):

    @classmethod
    def associated_class(cls) -> Type[DSModuleBase]:
        return DSEmbeddingBase
