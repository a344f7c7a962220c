"""Head bookkeeping for tensor-parallel attention (reference ``model_implementations/sharding/attn.py``)."""
from typing import Optional, Tuple


def get_local_heads(shard_rank: int, num_shards: int, n_heads_q: int, n_heads_kv: Optional[int] = None) -> Tuple[int, int]:
    """(local query heads, local kv heads).  MHA/GQA with enough KV heads: both split evenly (remainders to low ranks);
    fewer KV heads than ranks: each KV head is replicated over ``num_shards / n_heads_kv`` ranks."""
    if n_heads_q < num_shards:
        raise ValueError("There must be at least as many attention heads as there are shards.")
    if n_heads_kv is None or n_heads_kv == n_heads_q:
        base, extra = divmod(n_heads_q, num_shards)
        n = base + (1 if shard_rank < extra else 0)
        return n, n
    if n_heads_kv >= num_shards:
        if n_heads_kv % num_shards != 0 or n_heads_q % n_heads_kv != 0:
            raise ValueError("GQA sharding needs kv heads divisible by the shard count and q heads by kv heads")
        return n_heads_q // num_shards, n_heads_kv // num_shards
    if num_shards % n_heads_kv != 0 or n_heads_q % num_shards != 0:
        raise ValueError("With fewer KV heads than shards the shard count must be a multiple of the KV heads")
    return n_heads_q // num_shards, 1
