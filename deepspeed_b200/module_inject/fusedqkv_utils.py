"""Tensor-parallel slicing of fused QKV weights whose internal layout differs per model family (reference
``module_inject/fusedqkv_utils.py``)."""
import torch

from .tp_shard import get_num_kv_heads, get_shard_size_list

# how each family lays out its fused qkv along the output dim
#   "glmtype":    [q | k | v] blocks                      (chatglm, phi3 qkv_proj)
#   "bloomtype":  [heads, 3, head_dim] per-head interleave (bloom, gpt-neox, falcon-7b-style, megatron v2)
#   "codegentype": mp_num groups of [q | v | k]            (codegen)
#   "bigcodetype": multi-query: [q (all heads) | k | v] with ONE kv head (gpt_bigcode / starcoder)
_FUSED_QKV_TYPE = {"CodeGenBlock": "codegentype", "BloomBlock": "bloomtype", "GLMBlock": "glmtype", "MPTBlock": "glmtype",
                   "MptBlock": "glmtype", "BaichuanLayer": "glmtype", "QWenBlock": "qwentype", "FalconDecoderLayer": "bloomtype",
                   "GPTBigCodeBlock": "bigcodetype", "DecoderLayer": "glmtype", "Phi3DecoderLayer": "phi3type",
                   "GPTNeoXLayer": "bloomtype"}


def fused_type_of(module_or_name):
    name = module_or_name if isinstance(module_or_name, str) else type(module_or_name).__name__
    return _FUSED_QKV_TYPE.get(name)


def require_tp_fused_qkvw(name, mp_size):
    """Is ``name`` a fused qkv projection that needs layout-aware slicing?"""
    return mp_size > 1 and any(k in name for k in ("qkv_proj", "query_key_value", "attn.Wqkv", "self_attn.W_pack", "c_attn"))


def split_by_qkvlist_and_refuse(qkv_list, split_size, split_dim=0, cat_dim=0):
    """Split q, k and v each into TP pieces and re-fuse piece i of each: returns one tensor per rank."""
    pieces = [torch.split(t, split_size, dim=split_dim) if isinstance(split_size, int) else torch.split(t, split_size, dim=split_dim)
              for t in qkv_list]
    return [torch.cat([p[i] for p in pieces], dim=cat_dim) for i in range(len(pieces[0]))]


def prepare_tp_fused_qkvw(module, src, mp_size, gpu_index):
    """This rank's slice of a fused qkv weight or bias ``src`` (output dim first), whatever the family's layout."""
    if src is None:
        return None
    kind = fused_type_of(module) if not isinstance(module, str) else module
    n = src.shape[0]
    if kind in ("glmtype", "qwentype", "phi3type", None):
        kv = get_num_kv_heads()
        if kind == "phi3type" and kv is not None:
            # [q (H heads) | k (KV heads) | v (KV heads)] with H != KV: split each block separately
            head_dim_total = n
            # q rows : k rows : v rows = H : KV : KV  ->  need H; infer from module
            H = getattr(getattr(module, "self_attn", None), "num_heads", None) or getattr(
                getattr(getattr(module, "self_attn", None), "config", None), "num_attention_heads", None)
            if H is not None:
                d = head_dim_total // (H + 2 * kv)
                q, k, v = src[:H * d], src[H * d:(H + kv) * d], src[(H + kv) * d:]
                return torch.cat([q.chunk(mp_size, 0)[gpu_index], k.chunk(mp_size, 0)[gpu_index], v.chunk(mp_size, 0)[gpu_index]], 0)
        q, k, v = src.chunk(3, dim=0)
        sizes = get_shard_size_list(q.shape[0], mp_size)
        return split_by_qkvlist_and_refuse([q, k, v], sizes)[gpu_index]
    if kind == "bloomtype":
        # per-head [3, d] groups: whole heads go to a rank, internal order untouched
        sizes = get_shard_size_list(n, mp_size)
        return torch.split(src, sizes, dim=0)[gpu_index]
    if kind == "codegentype":
        mp_num = 4  # CodeGen checkpoints were trained with 4-way TPU model parallelism: 4 groups of [q | v | k]
        g = src.reshape(mp_num, 3, n // (3 * mp_num), *src.shape[1:])
        per = g.shape[2] // mp_size
        return g[:, :, gpu_index * per:(gpu_index + 1) * per].reshape(-1, *src.shape[1:])
    if kind == "bigcodetype":
        # multi-query attention: split q over ranks, replicate the single k/v head
        kv_dim = (n - _bigcode_q_dim(module, n)) // 2
        q, kvp = src[:n - 2 * kv_dim], src[n - 2 * kv_dim:]
        return torch.cat([q.chunk(mp_size, 0)[gpu_index], kvp], dim=0)
    raise ValueError(f"unknown fused qkv layout {kind}")


def _bigcode_q_dim(module, n):
    a = getattr(module, "attn", None)
    return getattr(a, "embed_dim", None) or (n * 0 + getattr(getattr(a, "c_attn", None), "in_features", n // 3))


def shard_value_with_share_qk(weight, bias, rank, world_size, shared_qk=True):
    """Layers where q and k share one projection ([qk | v]): split both halves."""
    def one(t):
        if t is None:
            return None
        a, b = t.chunk(2, dim=0)
        return torch.cat([a.chunk(world_size, 0)[rank], b.chunk(world_size, 0)[rank]], 0)
    return one(weight), one(bias)


def shard_chunk_mlp(weight, bias, rank, world_size):
    """Fused gate_up projection ([gate | up]): split each half over the ranks and re-fuse."""
    def one(t):
        if t is None:
            return None
        g, u = t.chunk(2, dim=0)
        return torch.cat([g.chunk(world_size, 0)[rank], u.chunk(world_size, 0)[rank]], 0)
    return one(weight), one(bias)
