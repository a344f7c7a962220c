"""Slice the token embedding along the hidden dimension (reference ``model_implementations/sharding/embedding.py``)."""
import torch

from .types import ShardingType
from .utils import get_shard_endpoints, shard_param


def shard_embedding_param(param: torch.Tensor, shard_rank: int, num_shards: int) -> torch.Tensor:
    return shard_param(param, ShardingType.INNER_DIMENSION, shard_rank, num_shards)


def sharded_embedding_dim(embedding_size: int, shard_rank: int, num_shards: int) -> int:
    g = 32 if embedding_size % 32 == 0 else 1
    s, e = get_shard_endpoints(embedding_size, shard_rank, num_shards, g)
    return e - s
