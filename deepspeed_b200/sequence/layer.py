"""DeepSpeed-Ulysses: all-to-all that trades the sequence shard for a head shard around attention.

Parity target: reference ``sequence/layer.py`` (``single_all_to_all :221``, ``_SeqAllToAll :257``,
``uneven_heads_all2all :111``, ``DistributedAttention :311``).  Every rank holds ``[s/P, b, h, d]`` (or batch
first); the first all-to-all produces ``[s, b, h/P, d]`` so any local attention kernel sees the full sequence
for its heads; the inverse restores the sequence sharding.  Heads not divisible by ``P`` are handled by giving
the first ``h % P`` ranks one extra head (variable split sizes), GQA-aware through ``num_kv_heads``.
"""
import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.utils import groups

_num_kv_heads = None


def set_num_kv_heads(n):
    global _num_kv_heads
    _num_kv_heads = n


def get_num_kv_heads():
    return _num_kv_heads


def _head_splits(h, world):
    base, rem = divmod(h, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


_gather_size_cache = {}


def _peer_gather_sizes(x, gather_idx, group):
    """Size of ``gather_idx`` on every rank (uneven head counts make them differ).  One tiny all-gather per
    distinct shape signature, then cached."""
    key = (id(group), gather_idx, tuple(x.shape), x.dtype)
    got = _gather_size_cache.get(key)
    if got is None:
        world = dist.get_world_size(group)
        dev = x.device if x.is_cuda else torch.device("cpu")
        mine = torch.tensor([x.shape[gather_idx]], dtype=torch.int64, device=dev)
        allv = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allv, mine, group=group)
        got = [int(v) for v in allv.tolist()]
        _gather_size_cache[key] = got
    return got


def _a2a_uneven(x, group, scatter_idx, gather_idx):
    """General all-to-all: scatter dim split as evenly as possible (first ranks get the remainder), gather dim
    concatenated from per-rank sizes that may differ."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    splits = _head_splits(x.shape[scatter_idx], world)
    gsizes = _peer_gather_sizes(x, gather_idx, group)
    ins = [p.contiguous() for p in torch.split(x, splits, dim=scatter_idx)]
    outs = []
    for p in range(world):
        shp = list(x.shape)
        shp[scatter_idx] = splits[rank]
        shp[gather_idx] = gsizes[p]
        outs.append(torch.empty(shp, dtype=x.dtype, device=x.device))
    dist.all_to_all(outs, ins, group=group)
    return torch.cat(outs, dim=gather_idx)


def single_all_to_all(input, scatter_idx, gather_idx, batch_dim_idx, group, async_op=False, handle=None, type=None):
    """Scatter ``scatter_idx`` (heads or sequence) over the group and gather ``gather_idx``."""
    world = dist.get_world_size(group)
    if world == 1:
        return input
    n = input.shape[scatter_idx]
    gsizes = _peer_gather_sizes(input, gather_idx, group)
    if n % world != 0 or len(set(gsizes)) != 1:
        return _a2a_uneven(input, group, scatter_idx, gather_idx)
    # even case: one all_to_all_single on a [P, ...] leading dimension
    x = input.reshape(list(input.shape[:scatter_idx]) + [world, n // world] + list(input.shape[scatter_idx + 1:]))
    x = x.movedim(scatter_idx, 0).contiguous()  # [P, ..., n/P, ...]
    out = torch.empty_like(x)
    dist.all_to_all_single(out, x, group=group)
    # out[p] = peer p's local chunk of the gather dimension for our scatter share -> concatenate along gather_idx
    return torch.cat(list(out.unbind(0)), dim=gather_idx)


class _SeqAllToAll(torch.autograd.Function):

    @staticmethod
    def forward(ctx, group, input, scatter_idx, gather_idx, batch_dim_idx=0, stream=None, handle=None, type=None,
                is_fwd=True):
        ctx.group, ctx.scatter_idx, ctx.gather_idx, ctx.batch_dim_idx = group, scatter_idx, gather_idx, batch_dim_idx
        return single_all_to_all(input, scatter_idx, gather_idx, batch_dim_idx, group)

    @staticmethod
    def backward(ctx, *grad_output):
        return (None, _SeqAllToAll.apply(ctx.group, grad_output[0], ctx.gather_idx, ctx.scatter_idx, ctx.batch_dim_idx),
                None, None, None, None, None, None, None)


class DistributedAttention(torch.nn.Module):
    """Wrap any local attention so it runs sequence-parallel (reference :311).

    ``local_attention(q, k, v, *args)`` receives full-sequence tensors with ``heads / P`` heads.
    ``scatter_idx`` is the head dimension, ``gather_idx`` the sequence dimension of the q/k/v layout.
    """

    def __init__(self, local_attention, sequence_process_group=None, scatter_idx: int = 2, gather_idx: int = 0,
                 sp_stream=None):
        super().__init__()
        self.local_attn = local_attention
        self.spg = sequence_process_group
        self.scatter_idx = scatter_idx
        self.gather_idx = gather_idx
        self.sp_overlap_comm = sp_stream is not None
        self.sp_stream = sp_stream
        self.overlap_handles = None

    def _group(self):
        return self.spg if self.spg is not None else groups._get_sequence_parallel_group()

    def forward(self, query, key, value, batch_dim_idx=None, *args, **kwargs):
        g = self._group()
        if batch_dim_idx is None:
            batch_dim_idx = 1 if self.gather_idx == 0 else 0
        if self.sp_overlap_comm and torch.cuda.is_available():
            # q/k all-to-alls on the side stream overlap the v projection still running on the main stream
            cur = torch.cuda.current_stream()
            self.sp_stream.wait_stream(cur)
            with torch.cuda.stream(self.sp_stream):
                q = _SeqAllToAll.apply(g, query, self.scatter_idx, self.gather_idx, batch_dim_idx)
                k = _SeqAllToAll.apply(g, key, self.scatter_idx, self.gather_idx, batch_dim_idx)
            v = _SeqAllToAll.apply(g, value, self.scatter_idx, self.gather_idx, batch_dim_idx)
            cur.wait_stream(self.sp_stream)
        else:
            q = _SeqAllToAll.apply(g, query, self.scatter_idx, self.gather_idx, batch_dim_idx)
            k = _SeqAllToAll.apply(g, key, self.scatter_idx, self.gather_idx, batch_dim_idx)
            v = _SeqAllToAll.apply(g, value, self.scatter_idx, self.gather_idx, batch_dim_idx)
        ctx = self.local_attn(q, k, v, *args, **kwargs)
        return _SeqAllToAll.apply(g, ctx, self.gather_idx, self.scatter_idx, batch_dim_idx)


# ---- functional helpers kept for callers of the reference module (``sequence/layer.py:59-170``) -------------------------
def post_all2all(permute_idx, res_shape):
    """Closure applied to an all-to-all result: optional permute, then reshape to ``res_shape``."""

    def post_func(t):
        if permute_idx is not None:
            t = t.permute(permute_idx)
        return t.reshape(res_shape).contiguous()

    return post_func


def pre_all2all_fun(permute_idx, inp_shape, input):
    t = input.reshape(inp_shape)
    return (t.permute(permute_idx) if permute_idx is not None else t).contiguous()


def _rotate_half(x):
    """``[x1, x2] -> [-x2, x1]`` over the two halves of the last dimension."""
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(t, freqs_cos, freqs_sin):
    """NeoX-style rotary embedding on the leading ``freqs_cos.shape[-1]`` features of ``t`` ([seq, ..., dim]); the rest
    passes through."""
    rot = freqs_cos.shape[-1]
    head, tail = t[..., :rot], t[..., rot:]
    head = head * freqs_cos + _rotate_half(head) * freqs_sin
    return head if tail.shape[-1] == 0 else torch.cat((head, tail), dim=-1)


def uneven_heads_all2all(input, scatter_idx, gather_idx, batch_dim_idx, group):
    """All-to-all for head counts that do not divide the sequence-parallel degree (first ranks take one extra head)."""
    if dist.get_world_size(group) == 1:
        return input
    return _a2a_uneven(input, group, scatter_idx, gather_idx)
