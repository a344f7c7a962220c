"""Checkpoint-engine protocol of the ragged engine (reference ``inference/v2/checkpoint/base_engine.py``): anything that
can enumerate ``(name, tensor)`` pairs can feed a model build."""
from abc import ABC, abstractmethod
from typing import Iterable, Tuple

import torch

MEGATRON = "megatron"
HUGGINGFACE = "huggingface"


class CheckpointEngineBase(ABC):

    @abstractmethod
    def parameters(self) -> Iterable[Tuple[str, torch.Tensor]]:
        """Yield every parameter once; tensors may be loaded lazily (one shard file at a time)."""
        ...

    def get(self, name: str):
        for n, t in self.parameters():
            if n == name:
                return t
        return None
