"""Enums + tiny helpers shared by the ragged engine (reference ``inference/v2/inference_utils.py``)."""
from enum import Enum, IntEnum

import torch


class NormTypeEnum(Enum):
    LayerNorm = "layer_norm"
    RMSNorm = "rms_norm"


class _Aliased(Enum):
    """Enum whose members accept several spellings; ``.value`` is the first (canonical) one."""

    def __new__(cls, *values):
        obj = object.__new__(cls)
        obj._value_ = values[0]
        for alias in values[1:]:
            cls._value2member_map_[alias] = obj
        obj._all_values = values
        return obj

    def __repr__(self):
        return f"<{type(self).__name__}.{self._name_}: {', '.join(map(repr, self._all_values))}>"


class DtypeEnum(_Aliased):
    fp16 = torch.float16, "torch.float16", "fp16", "float16", "half"
    fp32 = torch.float32, "torch.float32", "fp32", "float32", "float"
    bf16 = torch.bfloat16, "torch.bfloat16", "bf16", "bfloat16", "bfloat"
    int8 = torch.int8, "torch.int8", "int8"


class ActivationType(IntEnum):
    GELU = 0
    RELU = 1
    SILU = 2
    GEGLU = 3
    ReGLU = 4
    SiGLU = 5
    IDENTITY = 6
    InvalidType = -1


def is_gated(act_fn) -> bool:
    return ActivationType(act_fn) in (ActivationType.GEGLU, ActivationType.ReGLU, ActivationType.SiGLU)


def elem_size(dtype: torch.dtype) -> int:
    try:
        return torch.empty(0, dtype=dtype).element_size()
    except TypeError:
        raise ValueError(f"Unknown dtype size for {dtype}")


def ceil_div(a: int, b: int) -> int:
    return -(-a // b)
