"""Slice the attention output projection (row-parallel) (reference ``model_implementations/sharding/attn_out.py``)."""
from typing import Optional

import torch

from .attn import get_local_heads


def shard_attn_out_param(param: Optional[torch.Tensor], shard_rank: int, num_shards: int, head_size: int, n_heads_q: Optional[int] = None,
                         n_heads_kv: Optional[int] = None) -> Optional[torch.Tensor]:
    """Weight [out, hq * d] -> this rank's head columns; bias kept on rank 0 only."""
    if param is None or num_shards == 1:
        return param
    if param.dim() == 1:
        return param if shard_rank == 0 else None
    n_heads_q = n_heads_q if n_heads_q is not None else param.shape[1] // head_size
    start = sum(get_local_heads(r, num_shards, n_heads_q, n_heads_kv)[0] for r in range(shard_rank))
    mine = get_local_heads(shard_rank, num_shards, n_heads_q, n_heads_kv)[0]
    return param[:, start * head_size:(start + mine) * head_size]


def attn_out_in_features(out_features: int, shard_rank: int, num_shards: int, head_size: int, n_heads_q: Optional[int] = None,
                         n_heads_kv: Optional[int] = None) -> int:
    n_heads_q = n_heads_q if n_heads_q is not None else out_features // head_size
    return get_local_heads(shard_rank, num_shards, n_heads_q, n_heads_kv)[0] * head_size
