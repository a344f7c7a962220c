from .base_engine import CheckpointEngineBase  # noqa: F401
from .huggingface_engine import HuggingFaceCheckpointEngine  # noqa: F401
from .in_memory_engine import InMemoryModelEngine  # noqa: F401
