"""Tensor-parallel linear layers.

Parity target: reference ``module_inject/layers.py`` (``RowParallel :65``, ``ColumnParallel :97``,
``LinearAllreduce :300``, ``LinearLayer :370``, ``LmHeadLinearAllreduce``, fused-QKV / gate-up aware column
splits, uneven shard sizes, ``GatherReplacedLayerParams :238``).  Forward of a row-parallel layer ends in an
all-reduce, backward of a column-parallel layer's input gradient does; both go through
``deepspeed_b200.comm.inference_all_reduce`` which uses the one-shot NVLS kernel for small symmetric tensors.
"""
import torch
import torch.nn.functional as F
from torch import nn

from deepspeed_b200 import comm as dist


def shard_sizes(total, world, granularity=1):
    """Sizes of ``world`` contiguous shards of ``total`` in units of ``granularity`` (first shards larger)."""
    assert total % granularity == 0
    units = total // granularity
    base, rem = divmod(units, world)
    return [(base + (1 if r < rem else 0)) * granularity for r in range(world)]


def shard_bounds(total, world, rank, granularity=1):
    sizes = shard_sizes(total, world, granularity)
    start = sum(sizes[:rank])
    return start, start + sizes[rank]


class RowParallel(torch.autograd.Function):
    """Identity forward / all-reduce... inverse of ColumnParallel: all-reduce in forward, identity in backward."""

    @staticmethod
    def symbolic(graph, input):
        return input

    @staticmethod
    def forward(ctx, group, input, is_inference_mode=False):
        ctx.group = group
        if group is None or dist.get_world_size(group) == 1:
            return input
        input = input.contiguous()
        if is_inference_mode:
            dist.inference_all_reduce(input, group=group)
        else:
            dist.all_reduce(input, group=group)
        return input

    @staticmethod
    def backward(ctx, grad_output):
        return None, grad_output, None


class ColumnParallel(torch.autograd.Function):
    """Identity in forward, all-reduce of the input gradient in backward."""

    @staticmethod
    def forward(ctx, group, input):
        ctx.group = group
        return input

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.group is None or dist.get_world_size(ctx.group) == 1:
            return None, grad_output
        grad_output = grad_output.contiguous()
        dist.all_reduce(grad_output, group=ctx.group)
        return None, grad_output


class TensorParallel_Layer(nn.Module):
    keep_module_on_host = False

    def __init__(self, mp_group, name=None):
        super().__init__()
        self.mp_group = mp_group
        self.tp_world_size = dist.get_world_size(mp_group) if mp_group is not None else 1
        self.tp_index = dist.get_rank(mp_group) if mp_group is not None else 0
        self.name = name
        self.support_training = False

    def is_training_mode(self):
        return self.training

    def __deepcopy__(self, memo):
        """Process groups cannot be copied (or pickled): the clone shares ``mp_group`` (reference layers.py:214)."""
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        if self.mp_group is not None:
            memo[id(self.mp_group)] = self.mp_group
        for k, v in self.__dict__.items():
            new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _mark(self, *params):
        for p in params:
            if p is not None:
                p.tensor_model_parallel = True
                p.model_parallel = True
                p.ds_tp_world = self.tp_world_size
                # reference ``config_tp_params`` (layers.py:187): the parameter itself can be reassembled / re-split
                p.ds_is_replaced_module = True
                p.gather_params = self.gather_params
                p._tp_partition = self._tp_partition

    # ---- full <-> shard (reference ``gather_params`` / ``_tp_partition``; export, inspection, consolidation) -----------
    def _full_of(self, idx, shard):
        """Full tensor of parameter ``idx`` (0 = weight, 1 = bias) from this rank's shard; ``None`` = not sharded."""
        return None

    @torch.no_grad()
    def gather_params(self, params_list):
        """Replace every sharded parameter of ``params_list`` (``[weight, bias]`` order) by its full tensor; the shard is
        kept in ``param.data_partition`` for :meth:`_tp_partition`. Collective over the TP group."""
        for idx, p in enumerate(params_list):
            if p is None or self.tp_world_size == 1 or getattr(p, "data_partition", None) is not None:
                continue
            full = self._full_of(idx, p.data)
            if full is not None:
                p.data_partition = p.data
                p.data = full

    @torch.no_grad()
    def _tp_partition(self, params_list):
        """Undo :meth:`gather_params` (values written to the full tensor meanwhile are re-split into the shard)."""
        for idx, p in enumerate(params_list):
            part = getattr(p, "data_partition", None) if p is not None else None
            if part is None:
                continue
            part.copy_(self._shard_of(idx, p.data))
            p.data = part
            p.data_partition = None

    def _shard_of(self, idx, full):
        raise NotImplementedError


class LinearLayer(TensorParallel_Layer):
    """Column-parallel: each rank owns ``out/P`` output features (``weight`` rows)."""

    def __init__(self, module: nn.Module = None, mp_group=None, skip_partition=False, weight=None, bias=None,
                 granularity=1, fused_parts=1, **kwargs):
        super().__init__(mp_group, kwargs.get("name"))
        w = module.weight if module is not None else weight
        b = module.bias if module is not None and getattr(module, "bias", None) is not None else bias
        self.support_training = True
        self._fused_parts, self._granularity = fused_parts, granularity
        self._sharded = not (skip_partition or self.tp_world_size == 1)
        if skip_partition or self.tp_world_size == 1:
            self.weight, self.bias = nn.Parameter(w.data), (nn.Parameter(b.data) if b is not None else None)
        else:
            self.weight = nn.Parameter(_split_rows(w.data, self.tp_world_size, self.tp_index, fused_parts, granularity))
            self.bias = nn.Parameter(_split_rows(b.data, self.tp_world_size, self.tp_index, fused_parts,
                                                 granularity)) if b is not None else None
        self._mark(self.weight, self.bias)

    def forward(self, input):
        x = ColumnParallel.apply(self.mp_group, input) if self.training else input
        return F.linear(x, self.weight, self.bias)

    def _full_of(self, idx, shard):
        if not self._sharded:
            return None
        shards = _gather_list(shard, self.mp_group, 0)
        return _unsplit_rows(shards, self._fused_parts, self._granularity)

    def _shard_of(self, idx, full):
        return _split_rows(full, self.tp_world_size, self.tp_index, self._fused_parts, self._granularity)

    def extra_repr(self):
        return f"in={self.weight.shape[1]}, out_local={self.weight.shape[0]}, tp={self.tp_world_size}"


def _split_rows(t, world, rank, fused_parts=1, granularity=1):
    """Shard dim 0.  ``fused_parts`` > 1: the tensor is a concatenation (q|k|v or gate|up) and every part is
    sharded separately so each rank gets matching slices of all parts."""
    if fused_parts == 1:
        s, e = shard_bounds(t.shape[0], world, rank, granularity)
        return t[s:e].clone()
    parts = t.chunk(fused_parts, dim=0) if not isinstance(fused_parts, (list, tuple)) else torch.split(t, list(fused_parts), 0)
    outs = []
    for p in parts:
        s, e = shard_bounds(p.shape[0], world, rank, granularity)
        outs.append(p[s:e])
    return torch.cat(outs, dim=0).clone()


def _unsplit_rows(shards, fused_parts=1, granularity=1):
    """Inverse of :func:`_split_rows` given every rank's shard."""
    world = len(shards)
    if fused_parts == 1:
        return torch.cat(shards, dim=0)
    total = sum(sh.shape[0] for sh in shards)
    sizes = list(fused_parts) if isinstance(fused_parts, (list, tuple)) else [total // fused_parts] * fused_parts
    pieces = [[] for _ in sizes]
    for r, sh in enumerate(shards):
        off = 0
        for j, n in enumerate(sizes):
            a, b = shard_bounds(n, world, r, granularity)
            pieces[j].append(sh[off:off + (b - a)])
            off += b - a
    return torch.cat([torch.cat(ps, dim=0) for ps in pieces], dim=0)


class LinearAllreduce(TensorParallel_Layer):
    """Row-parallel: each rank owns ``in/P`` input features (``weight`` columns); outputs are summed."""

    def __init__(self, module: nn.Module = None, mp_group=None, weight=None, bias=None, granularity=1, **kwargs):
        super().__init__(mp_group, kwargs.get("name"))
        w = module.weight if module is not None else weight
        b = module.bias if module is not None and getattr(module, "bias", None) is not None else bias
        self.support_training = True
        if self.tp_world_size == 1:
            self.weight = nn.Parameter(w.data)
        else:
            s, e = shard_bounds(w.shape[1], self.tp_world_size, self.tp_index, granularity)
            self.weight = nn.Parameter(w.data[:, s:e].clone())
        self.bias = nn.Parameter(b.data) if b is not None else None  # replicated, added after the reduce
        self._granularity = granularity
        self._mark(self.weight)

    def forward(self, input):
        out = torch.matmul(input, self.weight.transpose(-1, -2))
        out = RowParallel.apply(self.mp_group, out, not self.training)
        if self.bias is not None:
            out = out + self.bias
        return out

    def _full_of(self, idx, shard):
        if idx > 0 or self.tp_world_size == 1:  # the bias is replicated
            return None
        return torch.cat(_gather_list(shard, self.mp_group, 1), dim=1)

    def _shard_of(self, idx, full):
        a, b = shard_bounds(full.shape[1], self.tp_world_size, self.tp_index, self._granularity)
        return full[:, a:b]


class LmHeadLinearAllreduce(LinearAllreduce):
    """LM head whose *input* (hidden) dimension is sharded: slices the incoming hidden states itself."""

    def forward(self, input):
        s, e = shard_bounds(input.shape[-1], self.tp_world_size, self.tp_index)
        if self.training:  # every rank only produces the gradient of its own slice of the hidden states
            input = ColumnParallel.apply(self.mp_group, input)
        out = torch.matmul(input[..., s:e], self.weight.transpose(-1, -2))
        out = RowParallel.apply(self.mp_group, out, not self.training)
        if self.bias is not None:
            out = out + self.bias
        return out


class GatherReplacedLayerParams:
    """Context manager: temporarily reassemble the full parameters of tensor-parallel layers (export / inspection;
    reference layers.py:238). ``module`` may be one TP layer or a container of them; collective over the TP group."""

    def __init__(self, params, module, enabled=True):
        self.enabled = enabled
        self.module = module
        self.params = list(params) if isinstance(params, (list, tuple)) or not torch.is_tensor(params) else [params]
        self._layers = []

    def __enter__(self):
        if not self.enabled:
            return self
        for m in self.module.modules():
            if isinstance(m, TensorParallel_Layer) and m.tp_world_size > 1:
                ps = [getattr(m, "weight", None), getattr(m, "bias", None)]
                m.gather_params(ps)
                self._layers.append((m, ps))
        return self

    def __exit__(self, *exc):
        for m, ps in self._layers:
            m._tp_partition(ps)
        self._layers.clear()
        return False


def _gather_list(t, group, dim):
    """Every TP rank's ``t`` (sizes along ``dim`` may differ: uneven head / granularity shards)."""
    world = dist.get_world_size(group)
    mine = torch.tensor([t.shape[dim]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    outs = []
    for n in sizes:
        shp = list(t.shape)
        shp[dim] = int(n.item())
        outs.append(torch.empty(shp, dtype=t.dtype, device=t.device))
    if len({tuple(o.shape) for o in outs}) == 1:
        dist.all_gather(outs, t.contiguous(), group=group)
    else:
        me = dist.get_rank(group)
        for i, o in enumerate(outs):
            if i == me:
                o.copy_(t)
            dist.broadcast(o, src=dist.get_global_rank(group, i) if group is not None else i, group=group)
    return outs


def _gather_dim(t, group, dim):
    return torch.cat(_gather_list(t, group, dim), dim=dim)


AUTOTP_TRAINING_MODE = False  # set through ``set_autotp_mode``: training builds autograd-aware TP layers


class EmbeddingLayer(nn.Module):
    """Plain embedding holding its weight as a parameter (reference ``module_inject/layers.py:EmbeddingLayer``)."""

    def __init__(self, weight_shape=None, dtype=torch.half, weight=None, bias=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(weight_shape, dtype=dtype) if weight is None else weight)

    def forward(self, input):
        return F.embedding(input, self.weight)


class Normalize(nn.Module):
    """LayerNorm whose weight / bias can be injected tensors (reference ``module_inject/layers.py:Normalize``)."""

    def __init__(self, dim=None, dtype=torch.float, eps=1e-5, weight=None, bias=None):
        super().__init__()
        if weight is not None:
            self.weight, self.bias = weight, bias
        else:
            self.norm = nn.LayerNorm(dim, eps=eps).to(dtype)
            self.weight, self.bias = self.norm.weight, self.norm.bias
        self.eps = eps

    def forward(self, input):
        return F.layer_norm(input, input.shape[-1:], self.weight, self.bias, self.eps)


# =====================================================================================================================
# Layout-specific variants (reference ``module_inject/layers.py``: fused_LinearLayer, conv_LinearLayer, GateUpPack_LinearLayer,
# Yuan_*, TensorParallelConv2d, ...).  They differ from LinearLayer / LinearAllreduce only in HOW the weight is cut.
# =====================================================================================================================
def get_auto_tp_mode():
    from deepspeed_b200.runtime.tensor_parallel.config import AUTOTP_MODE
    return AUTOTP_MODE.TRAINING if AUTOTP_TRAINING_MODE else AUTOTP_MODE.INFERENCE


def is_autotp_training_mode():
    return bool(AUTOTP_TRAINING_MODE)


class _PresplitColumn(TensorParallel_Layer):
    """Column-parallel layer built from an already computed local weight / bias."""

    def __init__(self, weight, bias, mp_group, name=None):
        super().__init__(mp_group, name)
        self.support_training = True
        self.weight = nn.Parameter(weight.contiguous())
        self.bias = nn.Parameter(bias.contiguous()) if bias is not None else None
        self._mark(self.weight, self.bias)

    def forward(self, input):
        x = ColumnParallel.apply(self.mp_group, input) if self.training else input
        return F.linear(x, self.weight, self.bias)


class fused_LinearLayer(_PresplitColumn):
    """Column-parallel layer over a FUSED q/k/v projection whose internal layout is family specific (per-head
    interleaved, grouped, blocked ...): the slice keeps whole heads together (``fusedqkv_utils.prepare_tp_fused_qkvw``)."""

    def __init__(self, module, mp_group, skip_partition=False, fused_module=None, **kwargs):
        from .fusedqkv_utils import prepare_tp_fused_qkvw
        world = dist.get_world_size(mp_group) if mp_group is not None else 1
        rank = dist.get_rank(mp_group) if mp_group is not None else 0
        w, b = module.weight.data, (module.bias.data if getattr(module, "bias", None) is not None else None)
        if not skip_partition and world > 1:
            owner = fused_module if fused_module is not None else kwargs.get("fused_type", "glmtype")
            w = prepare_tp_fused_qkvw(owner, w, world, rank)
            b = prepare_tp_fused_qkvw(owner, b, world, rank) if b is not None else None
        super().__init__(w.clone(), None if b is None else b.clone(), mp_group, kwargs.get("name"))


class GateUpPack_LinearLayer(_PresplitColumn):
    """Column-parallel layer over a packed ``[gate | up]`` projection (phi-3 ``gate_up_proj``): both halves are cut."""

    def __init__(self, module, mp_group, skip_partition=False, **kwargs):
        from .fusedqkv_utils import shard_chunk_mlp
        world = dist.get_world_size(mp_group) if mp_group is not None else 1
        rank = dist.get_rank(mp_group) if mp_group is not None else 0
        w, b = module.weight.data, (module.bias.data if getattr(module, "bias", None) is not None else None)
        if not skip_partition and world > 1:
            w, b = shard_chunk_mlp(w, b, rank, world)
        super().__init__(w.clone(), None if b is None else b.clone(), mp_group, kwargs.get("name"))


class conv_LinearLayer(_PresplitColumn):
    """Column-parallel layer built from a GPT-2 style ``Conv1D`` (weight stored ``[in, out]``)."""

    def __init__(self, module, mp_group, skip_partition=False, fused_parts=1, **kwargs):
        world = dist.get_world_size(mp_group) if mp_group is not None else 1
        rank = dist.get_rank(mp_group) if mp_group is not None else 0
        w = module.weight.data.t()  # -> [out, in]
        b = module.bias.data if getattr(module, "bias", None) is not None else None
        if not skip_partition and world > 1:
            w = _split_rows(w, world, rank, fused_parts)
            b = _split_rows(b, world, rank, fused_parts) if b is not None else None
        super().__init__(w.clone(), None if b is None else b.clone(), mp_group, kwargs.get("name"))


class Conv_LinearALlreduce(LinearAllreduce):
    """Row-parallel layer built from a ``Conv1D`` (weight ``[in, out]``)."""

    def __init__(self, module, mp_group, **kwargs):
        super().__init__(None, mp_group, weight=module.weight.data.t(), bias=getattr(module, "bias", None), **kwargs)


class Yuan_LinearLayer(_PresplitColumn):
    """Yuan's q/k projection stores ``[q | k]`` stacked: each half is cut separately so heads stay aligned."""

    def __init__(self, module, mp_group, skip_partition=False, **kwargs):
        from .fusedqkv_utils import shard_value_with_share_qk
        world = dist.get_world_size(mp_group) if mp_group is not None else 1
        rank = dist.get_rank(mp_group) if mp_group is not None else 0
        w, b = module.weight.data, (module.bias.data if getattr(module, "bias", None) is not None else None)
        if not skip_partition and world > 1:
            w, b = shard_value_with_share_qk(w, b, rank, world, True)
        super().__init__(w.clone(), None if b is None else b.clone(), mp_group, kwargs.get("name"))


class Yuan_LinearAllreduce(LinearAllreduce):
    """Row-parallel output projection of Yuan (input columns follow the same head split as the value projection)."""


class FusedModuleWrapper:
    """Wrap a fused module so attribute access is forwarded and ``forward`` splits the fused output when needed."""

    def __init__(self, fused_module: nn.Module):
        self.fused_module = fused_module

    def __getattr__(self, item):
        return getattr(self.__dict__["fused_module"], item)

    def __call__(self, *args, **kwargs):
        return self.fused_module(*args, **kwargs)


class RMSNormalize(nn.Module):
    """RMSNorm with injectable weight (reference ``RMSNormalize``); kernel path through ``transformer_ops.rms_norm``."""

    def __init__(self, dim=None, dtype=torch.float, eps=1e-5, weight=None):
        super().__init__()
        self.weight = weight if weight is not None else nn.Parameter(torch.ones(dim, dtype=dtype))
        self.eps = eps

    def forward(self, hidden_states):
        from deepspeed_b200.ops.kernels.transformer_ops import rms_norm
        return rms_norm(hidden_states, self.weight, self.eps)


class OPTEmbedding(EmbeddingLayer):
    """OPT learned positions: position ids come from the attention mask and are shifted by 2."""

    def __init__(self, weight_shape=None, weight=None, bias=None):
        super().__init__(weight_shape, weight=weight)
        self.offset = 2

    def forward(self, attention_mask: torch.LongTensor, past_key_values_length: int = 0, position_ids: int = 0):
        attention_mask = attention_mask.long()
        positions = (torch.cumsum(attention_mask, dim=1).type_as(attention_mask) * attention_mask).long() - 1
        positions = positions[:, past_key_values_length:]
        return super().forward(positions + self.offset)


# ---- tensor-parallel convolutions ------------------------------------------------------------------------------------
class TensorParallelConv2d(nn.Module):

    def __init__(self, conv, rank, world_size, shard_by_oc):
        super().__init__()
        self.rank, self.world_size, self.shard_by_oc = rank, world_size, shard_by_oc
        self.shard_weights(conv)

    def shard_weights(self, conv):
        """Keep this rank's output channels (``shard_by_oc``) or input channels of ``conv`` in place."""
        if self.world_size == 1:
            return
        oc, ic = conv.weight.shape[:2]
        if self.shard_by_oc:
            s, e = shard_bounds(oc, self.world_size, self.rank)
            conv.weight.data = conv.weight.data[s:e].clone()
            if conv.bias is not None:
                conv.bias.data = conv.bias.data[s:e].clone()
            conv.out_channels = e - s
        else:
            s, e = shard_bounds(ic, self.world_size, self.rank)
            conv.weight.data = conv.weight.data[:, s:e].clone()
            if conv.bias is not None:  # the bias is added once: only rank 0 keeps it
                conv.bias.data = conv.bias.data if self.rank == 0 else torch.zeros_like(conv.bias.data)
            conv.in_channels = e - s
            self._ic_range = (s, e)


class TensorParallelOcShardConv2d(TensorParallelConv2d):
    """Output-channel sharded conv: the result is this rank's channel slice (feed an Ic-sharded conv next)."""

    def __init__(self, conv, rank, world_size):
        super().__init__(conv, rank, world_size, True)
        self.conv = conv

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return self.conv(input)


class TensorParallelIcShardConv2d(TensorParallelConv2d):
    """Input-channel sharded conv: consumes this rank's channel slice, partial sums are all-reduced."""

    def __init__(self, conv, rank, world_size, mp_group=None):
        super().__init__(conv, rank, world_size, False)
        self.conv, self.mp_group = conv, mp_group

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        if self.world_size > 1 and input.shape[1] != self.conv.in_channels:
            s, e = self._ic_range
            input = input[:, s:e]
        out = self.conv(input)
        if self.world_size > 1:
            dist.inference_all_reduce(out, group=self.mp_group)
        return out


def set_autotp_mode(training=False):
    """Select what AutoTP layers are built for: training (autograd-aware collectives) or inference."""
    global AUTOTP_TRAINING_MODE
    AUTOTP_TRAINING_MODE = bool(training)


def move(tensor, device, copy=True):
    """Materialise ``tensor`` on ``device``: meta tensors become uninitialised storage, real ones are copied so the (often
    much larger, un-sharded) source can be freed."""
    if tensor.is_meta:
        return torch.empty_like(tensor, device=device)
    return tensor.to(device, copy=copy)
