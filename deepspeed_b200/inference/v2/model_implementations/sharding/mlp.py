"""Slice MLP projections (reference ``model_implementations/sharding/mlp.py``)."""
from typing import Optional

import torch

from .types import DEFAULT_SHARD_GRANULARITY, ShardingType
from .utils import get_shard_endpoints, shard_param


def shard_mlp_1_param(param: Optional[torch.Tensor], shard_rank: int, num_shards: int, gated: bool = False,
                      is_moe: bool = False) -> Optional[torch.Tensor]:
    """First projection (column-parallel).  ``gated``: gate and up are stacked and split independently.  ``is_moe``: a
    leading expert dim."""
    bias_dims = 2 if is_moe else 1
    return shard_param(param, ShardingType.OUTER_DIMENSION, shard_rank, num_shards, num_concatenated_matrices=2 if gated else 1,
                       granularity=DEFAULT_SHARD_GRANULARITY, bias_dims=bias_dims)


def shard_mlp_2_param(param: Optional[torch.Tensor], shard_rank: int, num_shards: int, is_moe: bool = False) -> Optional[torch.Tensor]:
    """Second projection (row-parallel); its bias lives on rank 0."""
    bias_dims = 2 if is_moe else 1
    return shard_param(param, ShardingType.INNER_DIMENSION, shard_rank, num_shards, granularity=DEFAULT_SHARD_GRANULARITY,
                       bias_dims=bias_dims)


def sharded_intermediate_dim(intermediate_size: int, num_shards: int, shard_rank: int) -> int:
    g = DEFAULT_SHARD_GRANULARITY if intermediate_size % DEFAULT_SHARD_GRANULARITY == 0 else 1
    s, e = get_shard_endpoints(intermediate_size, shard_rank, num_shards, g)
    return e - s
