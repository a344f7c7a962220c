"""Abstract model protocol of the ragged engine (reference ``model_implementations/inference_model_base.py``): what the
engine needs from any model - KV-cache requirements, the forward over a ragged batch, and the parameter-transform hooks
the containers call."""
from abc import ABC, abstractmethod
from typing import Tuple

import torch


class DSInferenceModelBase(torch.nn.Module, ABC):

    @abstractmethod
    def get_kv_requirements(self, sequence, max_new_tokens: int, max_new_blocks: int) -> Tuple[int, int]:
        """(tokens that can be scheduled, KV blocks needed) for ``sequence``."""

    @abstractmethod
    def kv_cache_config(self, *a, **k):
        """KVCacheConfig describing the cache this model needs."""

    @abstractmethod
    def forward(self, wrapped_batch) -> torch.Tensor:
        """Logits of the last token of every sequence in the ragged batch."""

    def get_remaining_block_capacity(self, sequence) -> int:
        bs = self.kv_block_size if hasattr(self, "kv_block_size") else 128
        return (-sequence.seen_tokens) % bs

    def maybe_allocate_kv(self, sequence, n_new_tokens: int) -> None:
        _, n_blocks = self.get_kv_requirements(sequence, n_new_tokens, self.state_manager.free_blocks)
        if n_blocks > 0:
            sequence.extend_kv_cache(self.state_manager.allocate_blocks(n_blocks))

    def maybe_free_kv(self, sequence) -> None:
        """Dense caches never release blocks mid-sequence (sliding-window models would here)."""
        return None
