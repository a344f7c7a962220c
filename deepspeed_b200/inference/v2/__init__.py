"""Ragged / continuous-batching inference ("FastGen" role).  Reference: ``deepspeed/inference/v2``."""
from .config_v2 import RaggedInferenceEngineConfig, DeepSpeedTPConfig, DSStateManagerConfig, KVCacheConfig  # noqa: F401
from .engine_v2 import InferenceEngineV2  # noqa: F401
from .engine_factory import build_hf_engine, build_engine_from_model  # noqa: F401
from .scheduling_utils import SchedulingResult, SchedulingError  # noqa: F401
from .engine_factory import build_engine_from_ds_checkpoint  # noqa: F401,E402
