"""State-manager config models (reference ``inference/v2/ragged/manager_configs.py``); defined with the engine config in
``config_v2.py``."""
from ..config_v2 import AllocationMode, DSStateManagerConfig, KVCacheConfig, KVCacheType, MemoryConfig  # noqa: F401
