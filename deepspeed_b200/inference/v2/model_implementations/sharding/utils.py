"""Generic shard arithmetic (reference ``model_implementations/sharding/utils.py``)."""
from typing import Optional, Tuple

import torch

from .types import DEFAULT_SHARD_GRANULARITY, ShardingType


def get_shard_endpoints(dim_size: int, shard_rank: int, num_shards: int, granularity: int = DEFAULT_SHARD_GRANULARITY) -> Tuple[int, int]:
    """[start, end) of ``shard_rank``'s slice of a dimension: whole ``granularity`` blocks, remainder blocks to the
    lowest ranks."""
    assert dim_size % granularity == 0, f"Dimension size {dim_size} must be divisible by granularity {granularity}"
    blocks = dim_size // granularity
    base, extra = divmod(blocks, num_shards)
    start = shard_rank * base + min(shard_rank, extra)
    end = start + base + (1 if shard_rank < extra else 0)
    return start * granularity, end * granularity


def shard_param(param: Optional[torch.Tensor], shard_mode: ShardingType, shard_rank: int, num_shards: int, num_concatenated_matrices: int = 1,
                granularity: int = 32, bias_dims: int = 1) -> Optional[torch.Tensor]:
    """Slice ``param`` for ``shard_rank``.  ``num_concatenated_matrices``: the outer dim holds that many stacked matrices
    (fused gate/up, fused qkv with equal parts) that must each be split independently.  A bias (``dim == bias_dims``) of
    an INNER-sharded linear is kept whole on rank 0 and dropped elsewhere (it is added once after the all-reduce)."""
    if param is None:
        return None
    if num_shards == 1:
        return param
    is_bias = param.dim() == bias_dims
    if shard_mode == ShardingType.INNER_DIMENSION:
        if is_bias:
            return param if shard_rank == 0 else None
        s, e = get_shard_endpoints(param.shape[-1], shard_rank, num_shards, min(granularity, param.shape[-1] // num_shards) or 1)
        return param[..., s:e]
    # outer dimension
    axis = param.dim() - bias_dims if is_bias else param.dim() - 2
    total = param.shape[axis]
    assert total % num_concatenated_matrices == 0
    each = total // num_concatenated_matrices
    g = granularity if each % (granularity * 1) == 0 and each // granularity >= num_shards else 1
    pieces = []
    for m in range(num_concatenated_matrices):
        s, e = get_shard_endpoints(each, shard_rank, num_shards, g)
        pieces.append(param.narrow(axis, m * each + s, e - s))
    return torch.cat(pieces, dim=axis) if len(pieces) > 1 else pieces[0]
