"""Slice the LM head along the vocabulary (reference ``model_implementations/sharding/unembed.py``)."""
import torch

from .types import ShardingType
from .utils import get_shard_endpoints, shard_param


def shard_unembed_param(param: torch.Tensor, shard_rank: int, num_shards: int) -> torch.Tensor:
    return shard_param(param, ShardingType.OUTER_DIMENSION, shard_rank, num_shards, granularity=1)


def sharded_unembed_dim(vocab_size: int, shard_rank: int, num_shards: int) -> int:
    s, e = get_shard_endpoints(vocab_size, shard_rank, num_shards, 1)
    return e - s
