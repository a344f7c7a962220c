"""Configs of the optimised linear family (reference ``linear/config.py``)."""
from dataclasses import dataclass, field
from typing import List

import torch


@dataclass
class LoRAConfig:
    """``lora_r`` rank, ``lora_alpha`` scaling numerator, ``base_weight_sharding``: shard the frozen base weight
    this many ways over the DP world (gathered on use), ``offload``/``offload_ratio``: keep (part of) the base
    weight on the host, ``delay_lora_init``: create adapters later via ``init_lora``, ``target_mods``: module-name
    suffixes the ``Init`` context converts."""
    lora_r: int = 64
    lora_alpha: float = 16.0
    base_weight_sharding: int = 1
    offload: bool = False
    offload_ratio: float = 0.0
    delay_lora_init: bool = False
    target_mods: List[str] = field(
        default_factory=lambda: ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"])


@dataclass
class QuantizationConfig:
    """``q_bits`` storage bits (8 / 6 / 12 / 4), ``mantissa_bits`` of the FP format, ``group_size`` elements per
    scale, ``q_dtype`` container dtype."""
    q_bits: int = 8
    mantissa_bits: int = 3
    group_size: int = 512
    q_dtype: torch.dtype = torch.uint8
    q_range_dtype: torch.dtype = torch.float8_e4m3fn if hasattr(torch, "float8_e4m3fn") else torch.float16
