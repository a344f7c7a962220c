"""Ragged token embedding (reference ``modules/implementations/embedding/ragged_embedding.py``)."""
from typing import Any, Dict

import torch

from ....inference_utils import DtypeEnum
from ....kernels.ragged_ops import RaggedEmbeddingKernel
from ...configs import DSEmbeddingsConfig
from ...interfaces import DSEmbeddingBase, DSEmbeddingRegistry


@DSEmbeddingRegistry.register_module
class DSRaggedEmbedding(DSEmbeddingBase):

    @staticmethod
    def name() -> str:
        return "ragged_embedding"

    @staticmethod
    def supports_config(config: DSEmbeddingsConfig) -> bool:
        return not config.use_token_type and config.output_normalization is None

    def __init__(self, config: DSEmbeddingsConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)
        self.kernel = RaggedEmbeddingKernel(DtypeEnum(config.residual_dtype).value, torch.int32, config.embedding_dim)
        self._out = None

    @property
    def output(self) -> torch.Tensor:
        return self._out

    def forward(self, ragged_batch, word_embeddings, position_embeddings=None, token_type_ids=None, token_type_embeddings=None):
        ids = ragged_batch.input_ids()
        out = torch.empty(ids.numel(), self._config.embedding_dim, dtype=word_embeddings.dtype, device=word_embeddings.device)
        pos = ragged_batch.pos_of() if (self._config.positional_embedding and position_embeddings is not None) else None
        self._out = self.kernel(out, ids, word_embeddings, pos, position_embeddings if pos is not None else None,
                                self._config.positional_offset)
        return self._out
