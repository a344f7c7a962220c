"""Slice the fused ``[q | k | v]`` projection (reference ``model_implementations/sharding/qkv.py``)."""
from typing import Optional

import torch

from .attn import get_local_heads


def shard_qkv_param(param: Optional[torch.Tensor], shard_rank: int, num_shards: int, head_size: int, n_heads_q: Optional[int] = None,
                    n_heads_kv: Optional[int] = None) -> Optional[torch.Tensor]:
    """``param``: [(hq + 2 hkv) * d, in] weight or [(hq + 2 hkv) * d] bias.  Query heads are split contiguously; KV heads
    are split, or - when there are fewer of them than ranks - the owning head is replicated."""
    if param is None or num_shards == 1:
        return param
    if n_heads_q is None:
        assert param.shape[0] % (3 * head_size) == 0
        n_heads_q = n_heads_kv = param.shape[0] // (3 * head_size)
    n_heads_kv = n_heads_kv or n_heads_q
    d = head_size
    q, k, v = param[:n_heads_q * d], param[n_heads_q * d:(n_heads_q + n_heads_kv) * d], param[(n_heads_q + n_heads_kv) * d:]
    lq, lkv = get_local_heads(shard_rank, num_shards, n_heads_q, n_heads_kv)
    if n_heads_kv == n_heads_q:  # uneven splits allowed: prefix sums of the per-rank head counts
        start = sum(get_local_heads(r, num_shards, n_heads_q)[0] for r in range(shard_rank))
        sl = slice(start * d, (start + lq) * d)
        return torch.cat([q[sl], k[sl], v[sl]], dim=0)
    qs = q[shard_rank * lq * d:(shard_rank + 1) * lq * d]
    if n_heads_kv >= num_shards:
        ks = slice(shard_rank * lkv * d, (shard_rank + 1) * lkv * d)
    else:
        h = shard_rank // (num_shards // n_heads_kv)
        ks = slice(h * d, (h + 1) * d)
    return torch.cat([qs, k[ks], v[ks]], dim=0)


def qkv_out_features(in_features: int, shard_rank: int, num_shards: int, head_size: int, n_heads_q: Optional[int] = None,
                     n_heads_kv: Optional[int] = None) -> int:
    n_heads_q = n_heads_q if n_heads_q is not None else in_features // head_size
    lq, lkv = get_local_heads(shard_rank, num_shards, n_heads_q, n_heads_kv)
    return (lq + 2 * lkv) * head_size
