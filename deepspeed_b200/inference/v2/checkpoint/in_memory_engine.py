"""Checkpoint engine over a live ``torch.nn.Module`` (reference ``inference/v2/checkpoint/in_memory_engine.py``)."""
from typing import Iterable, Tuple

import torch

from .base_engine import CheckpointEngineBase


class InMemoryModelEngine(CheckpointEngineBase):
    """Yields the module's own parameters / persistent buffers (no copies); handy right after training (RLHF) or in tests."""

    def __init__(self, model: torch.nn.Module) -> None:
        super().__init__()
        self.model = model
        self.model_config = getattr(model, "config", None)

    def parameters(self) -> Iterable[Tuple[str, torch.Tensor]]:
        yield from self.model.state_dict().items()

    def get(self, name: str):
        return self.model.state_dict().get(name)
