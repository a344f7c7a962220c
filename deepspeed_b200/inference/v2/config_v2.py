"""Configs of the ragged engine (reference ``inference/v2/config_v2.py``, ``ragged/manager_configs.py``)."""
from enum import Enum
from typing import Optional

from pydantic import Field

from deepspeed_b200.runtime.config_utils import DeepSpeedConfigModel


class DeepSpeedTPConfig(DeepSpeedConfigModel):
    tp_size: int = 1


class QuantizationConfig(DeepSpeedConfigModel):
    quantization_mode: Optional[str] = None  # None | "wf6af16" | "fp8" | "int8"


class KVCacheType(Enum):
    DENSE = "dense"
    LOCAL = "local"


class AllocationMode(Enum):
    RESERVE = "reserve"      # leave `size` bytes free, use the rest of HBM for KV blocks
    ALLOCATE = "allocate"    # allocate exactly `size` blocks


class MemoryConfig(DeepSpeedConfigModel):
    mode: AllocationMode = AllocationMode.RESERVE
    size: int = Field(8_000_000_000, gt=0)


class KVCacheConfig(DeepSpeedConfigModel):
    block_size: int = 128
    num_allocation_groups: int = Field(1, gt=0)
    cache_shape: tuple = ()          # (layers, kv_heads, head_dim)
    cache_dtype: str = "bf16"
    max_blocks_per_allocation_group: int = 64


class DSStateManagerConfig(DeepSpeedConfigModel):
    max_tracked_sequences: int = Field(2048, gt=0)
    max_ragged_batch_size: int = Field(768, gt=0)
    max_ragged_sequence_count: int = Field(512, gt=0)
    max_context: int = Field(8192, gt=0)
    memory_config: MemoryConfig = MemoryConfig()
    offload: bool = False


class RaggedInferenceEngineConfig(DeepSpeedConfigModel):
    tensor_parallel: DeepSpeedTPConfig = Field({}, alias="tp")
    state_manager: DSStateManagerConfig = Field({}, alias="manager")
    quantization: QuantizationConfig = {}
    cuda_graph_decode: bool = True   # capture pure-decode steps per padded batch size
