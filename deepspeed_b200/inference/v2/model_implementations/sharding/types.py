"""How a parameter is split across tensor-parallel ranks (reference ``model_implementations/sharding/types.py``)."""
from enum import Enum

DEFAULT_SHARD_GRANULARITY = 32  # sliced dimensions stay multiples of this so vectorised kernels keep their alignment


class ShardingType(Enum):
    OUTER_DIMENSION = 0  # split rows (output features): column-parallel linear
    INNER_DIMENSION = 1  # split columns (input features): row-parallel linear
