"""Small enums shared by the inference stacks (reference ``utils/types.py``)."""
from enum import IntEnum

# values are part of the kernel ABI (passed as ints): keep the numbering
ActivationFuncType = IntEnum("ActivationFuncType", ["UNKNOWN", "GELU", "ReLU", "GATED_GELU", "GATED_SILU"], start=0)
NormType = IntEnum("NormType", ["UNKNOWN", "LayerNorm", "GroupNorm", "RMSNorm"], start=0)

GATED_ACTIVATION_TYPES = [a for a in ActivationFuncType if a.name.startswith("GATED_")]
