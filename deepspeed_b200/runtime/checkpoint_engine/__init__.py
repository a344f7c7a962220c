"""Checkpoint engines (reference: ``runtime/checkpoint_engine/*``): pluggable persistence back-ends."""
from .checkpoint_engine import CheckpointEngine  # noqa: F401
from .torch_checkpoint_engine import TorchCheckpointEngine  # noqa: F401
from .async_checkpoint_engine import AsyncCheckpointEngine  # noqa: F401
from .nebula_checkpoint_engine import NebulaCheckpointEngine  # noqa: F401


def build_checkpoint_engine(config):
    """Select the engine: async/"nebula"-style tiered writer when configured, else ``torch.save``."""
    writer = getattr(config.checkpoint_config, "writer", None)
    if getattr(config.nebula_config, "enabled", False) or (writer and writer.get("type", "").lower() == "nebula"):
        from .nebula_checkpoint_engine import NebulaCheckpointEngine
        return NebulaCheckpointEngine(config)
    if writer and writer.get("type", "").lower() in ("async", "fast"):
        return AsyncCheckpointEngine(config)
    return TorchCheckpointEngine(config)
