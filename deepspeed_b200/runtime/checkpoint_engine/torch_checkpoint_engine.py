"""Synchronous ``torch.save`` / ``torch.load`` engine (reference: ``torch_checkpoint_engine.py``)."""
import torch

from deepspeed_b200.utils.logging import logger
from .checkpoint_engine import CheckpointEngine


class TorchCheckpointEngine(CheckpointEngine):

    def create(self, tag):
        logger.debug(f"[Torch] Checkpoint {tag} is about to be saved!")

    def save(self, state_dict, path: str):
        tmp = f"{path}.tmp"
        torch.save(state_dict, tmp)
        import os
        os.replace(tmp, path)  # atomic publish: a crash never leaves a truncated shard

    def load(self, path: str, map_location=None):
        return torch.load(path, map_location=map_location, weights_only=False)

    def commit(self, tag):
        logger.debug(f"[Torch] Checkpoint {tag} is ready now!")
        return True
