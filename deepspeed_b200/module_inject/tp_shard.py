"""Global tensor-parallel sharding parameters + shard-size arithmetic (reference ``module_inject/tp_shard.py``)."""
_num_kv_heads = None
_n_embd = None
_tp_grain_size = 1


def set_num_kv_heads(num):
    global _num_kv_heads
    _num_kv_heads = num


def get_num_kv_heads():
    return _num_kv_heads


def set_num_attention_heads(num):
    global _num_attention_heads
    _num_attention_heads = num


def get_num_attention_heads():
    return globals().get("_num_attention_heads")


def set_n_embd(num):
    global _n_embd
    _n_embd = num


def get_n_embd():
    return _n_embd


def set_tp_grain_size(num):
    global _tp_grain_size
    _tp_grain_size = max(1, int(num))


def get_shard_size(total_size, mp_size, name=None, rank=None):
    """Size of ``rank``'s shard of a dimension of ``total_size``.  Attention-shaped dimensions are divided in whole KV
    heads, MLP / vocabulary dimensions in multiples of the TP grain; earlier ranks absorb the remainder."""
    if rank is None:
        from deepspeed_b200 import comm as dist
        rank = dist.get_rank() % mp_size if dist.is_initialized() else 0
    last_linear = ("lm_head", "embed_out")
    if _num_kv_heads is not None and total_size % _num_kv_heads == 0 and name not in last_linear and "mlp" not in str(name):
        units, unit = _num_kv_heads, total_size // _num_kv_heads
    elif total_size >= 64 and total_size % _tp_grain_size == 0:
        units, unit = total_size // _tp_grain_size, _tp_grain_size
    else:
        units, unit = total_size, 1
    return (units // mp_size + (1 if rank < units % mp_size else 0)) * unit


def get_shard_size_list(total_size, mp_size, name=None):
    return [get_shard_size(total_size, mp_size, name, r) for r in range(mp_size)]
