"""Local Hugging Face checkpoint reader (reference ``inference/v2/checkpoint/huggingface_engine.py``).  There is no hub
access here: ``model_name_or_path`` must be a directory with ``config.json`` and safetensors / torch shards."""
import os

from ..engine_factory import HuggingFaceCheckpointEngine as _Reader
from .base_engine import CheckpointEngineBase


class HuggingFaceCheckpointEngine(_Reader, CheckpointEngineBase):

    def __init__(self, model_name_or_path: str, auth_token: str = None, **hf_kwargs) -> None:
        if not os.path.isdir(model_name_or_path):
            raise FileNotFoundError(f"{model_name_or_path} is not a local directory (hub downloads are not available)")
        super().__init__(model_name_or_path)
        self.model_name_or_path = model_name_or_path
        self.auth_token = auth_token
