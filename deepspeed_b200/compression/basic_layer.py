"""Compressible drop-in layers (reference ``compression/basic_layer.py``): each layer keeps the dense weight and
applies the enabled techniques (sparse / row / head / channel pruning masks, weight + activation fake
quantisation with straight-through gradients) in ``forward``; ``fix_*_helper`` bakes a technique in for export."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .utils import AsymQuantizer, BinaryQuantizer, SymQuantizer, TernaryQuantizer, TopKBinarizer


class QuantAct(nn.Module):
    """Activation fake-quantiser with EMA range calibration (static) or per-call ranges (dynamic)."""

    def __init__(self, act_range_momentum=0.95, quant_mode="symmetric"):
        super().__init__()
        self.act_range_momentum = act_range_momentum
        self.quant_mode = quant_mode
        self.act_function = SymQuantizer.apply if quant_mode == "symmetric" else AsymQuantizer.apply
        self.register_buffer("x_min_max", torch.zeros(2))

    def forward(self, x, num_bits, *args):
        if self.training:
            mn, mx = x.detach().min(), x.detach().max()
            if self.x_min_max[0] == self.x_min_max[1]:
                self.x_min_max[0], self.x_min_max[1] = mn, mx
            self.x_min_max[0] = self.x_min_max[0] * self.act_range_momentum + mn * (1 - self.act_range_momentum)
            self.x_min_max[1] = self.x_min_max[1] * self.act_range_momentum + mx * (1 - self.act_range_momentum)
        return self.act_function(x, num_bits, self.x_min_max[0], self.x_min_max[1])


class _CompressMixin:
    """State + switches shared by every compressible layer."""

    def _init_compress_state(self):
        self.sparse_pruning_method = self.row_pruning_method = self.head_pruning_method = None
        self.channel_pruning_method = None
        self.activation_quantization_method = None
        self.weight.start_bits = None
        self.weight.target_bits = None
        self.weight.q_period = None
        self.weight_quantization_enabled_in_forward = False
        self.weight_quantization_enabled = False
        self.sparse_pruning_enabled = self.row_pruning_enabled = self.head_pruning_enabled = False
        self.channel_pruning_enabled = False
        self.activation_quantization_enabled = False
        self.weight_quantize_num_groups = 1

    # ---- weight quantisation
    def enable_weight_quantization(self, start_bits, target_bits, quantization_period, weight_quantization_enabled_in_forward,
                                   quantization_type, num_groups):
        self.weight.start_bits, self.weight.target_bits, self.weight.q_period = start_bits, target_bits, quantization_period
        self.weight_quantization_enabled_in_forward = weight_quantization_enabled_in_forward
        if self.weight_quantization_enabled_in_forward:
            if self.weight.target_bits >= 3:
                self.weight_quantizer = SymQuantizer.apply if quantization_type == "symmetric" else AsymQuantizer.apply
            elif self.weight.target_bits == 2:
                assert quantization_type == "symmetric", "Only symmetric quantization is supported for ternary weights"
                self.weight_quantizer = TernaryQuantizer.apply
            elif self.weight.target_bits == 1:
                assert quantization_type == "symmetric", "Only symmetric quantization is supported for binary weights"
                self.weight_quantizer = BinaryQuantizer.apply
            self.weight_quantize_num_groups = num_groups

    def fix_weight_quantization(self):
        self.weight.data = self.weight_quantizer(self.weight, self.weight.target_bits, None, None,
                                                 self.weight_quantize_num_groups).data
        self.weight_quantization_enabled_in_forward = False
        return None

    # ---- activation quantisation
    def enable_activation_quantization(self, bits, quantization_type, range_calibration):
        assert bits in (4, 8), "Only 4/8 bits activation quantization are supported for now"
        self.activation_quantization_bits = bits
        self.activation_quantization_method = f"{quantization_type}_{range_calibration}"
        if range_calibration == "static":
            self.activation_quantizer = QuantAct(quant_mode=quantization_type)
        else:
            self.activation_quantizer = SymQuantizer.apply if quantization_type == "symmetric" else AsymQuantizer.apply

    def _quant_input(self, x):
        if "dynamic" in self.activation_quantization_method:
            groups = x.numel() // x.size(-1)
            return self.activation_quantizer(x, self.activation_quantization_bits, None, None, groups)
        return self.activation_quantizer(x, self.activation_quantization_bits)

    # ---- sparse pruning
    def enable_sparse_pruning(self, ratio, method):
        self.sparse_pruning_ratio, self.sparse_pruning_method = ratio, method
        if method == "l1":
            self.register_buffer("sparse_pruning_mask", self._l1_mask(self.weight, ratio).to(self.weight.device))
        elif method == "topk":
            self.sparse_mask_scores = nn.Parameter(torch.empty_like(self.weight))
            nn.init.kaiming_uniform_(self.sparse_mask_scores, a=math.sqrt(5))
        else:
            raise NotImplementedError(f"sparse pruning method {method}")

    @staticmethod
    def _l1_mask(w, dense_ratio):
        k = int(w.numel() * (1 - dense_ratio))
        if k <= 0:
            return torch.ones_like(w, dtype=torch.bool)
        thr = w.detach().abs().flatten().kthvalue(k).values
        return w.detach().abs() > thr

    def get_mask(self, pruning_type="sparse"):
        if pruning_type == "sparse":
            if self.sparse_pruning_method == "l1":
                return self.sparse_pruning_mask.to(self.weight.device)
            return TopKBinarizer.apply(self.sparse_mask_scores, self.sparse_pruning_ratio, False)
        if pruning_type == "row":
            if self.row_pruning_method == "l1":
                return self.row_pruning_mask.to(self.weight.device)
            return TopKBinarizer.apply(self.row_mask_scores, self.row_pruning_ratio, False)
        if pruning_type == "head":
            return TopKBinarizer.apply(self.head_pruning_scores, self.head_pruning_ratio, False)
        if pruning_type == "channel":
            if self.channel_pruning_method == "l1":
                return self.channel_pruning_mask.to(self.weight.device)
            return TopKBinarizer.apply(self.channel_mask_scores, self.channel_pruning_ratio, False)
        raise NotImplementedError(pruning_type)

    def fix_sparse_pruning_helper(self):
        mask = self.get_mask("sparse")
        self.weight.data = self.weight.data * mask
        for a in ("sparse_pruning_mask", "sparse_mask_scores"):
            if hasattr(self, a):
                delattr(self, a)
        self.sparse_pruning_method = None
        self.sparse_pruning_enabled = False
        return None


class LinearLayer_Compress(nn.Linear, _CompressMixin):

    def __init__(self, *kargs, bias=True):
        super().__init__(*kargs, bias=bias)
        self._init_compress_state()

    def __repr__(self):
        return (f"LinearLayer_Compress(in_features={self.in_features}, out_features={self.out_features}, "
                f"bias={self.bias is not None}, sparse pruning={self.sparse_pruning_method}, row pruning="
                f"{self.row_pruning_method}, head pruning={self.head_pruning_method}, activation quantization="
                f"{self.activation_quantization_method}, weight_quantization={self.weight.target_bits})")

    def enable_row_pruning(self, ratio, method):
        self.row_pruning_ratio, self.row_pruning_method = ratio, method
        if method == "l1":
            norms = self.weight.detach().abs().sum(dim=1)
            k = int(norms.numel() * (1 - ratio))
            thr = norms.kthvalue(k).values if k > 0 else norms.min() - 1
            self.register_buffer("row_pruning_mask", (norms > thr).view(-1, 1))
        elif method == "topk":
            self.row_mask_scores = nn.Parameter(torch.empty(self.out_features, 1))
            nn.init.kaiming_uniform_(self.row_mask_scores, a=math.sqrt(5))
        else:
            raise NotImplementedError

    def enable_head_pruning(self, ratio, method, num_heads):
        assert self.in_features % num_heads == 0, "head pruning applies to the attention output projection"
        self.num_heads, self.head_pruning_ratio, self.head_pruning_method = num_heads, ratio, method
        if method != "topk":
            raise NotImplementedError("only topk head pruning is supported")
        self.head_pruning_scores = nn.Parameter(torch.empty(1, num_heads))
        nn.init.kaiming_uniform_(self.head_pruning_scores, a=math.sqrt(5))

    def fix_row_col_pruning_helper(self, mask=None, dim_reduction=False):
        if mask is None:
            mask = self.get_mask("row").bool()
            if dim_reduction:
                keep = mask.view(-1)
                self.weight = nn.Parameter(self.weight.data[keep])
                if self.bias is not None:
                    self.bias = nn.Parameter(self.bias.data[keep])
                self.out_features = int(keep.sum())
            else:
                self.weight.data = self.weight.data * mask.view(-1, 1)
                if self.bias is not None:
                    self.bias.data = self.bias.data * mask.view(-1)
            for a in ("row_pruning_mask", "row_mask_scores"):
                if hasattr(self, a):
                    delattr(self, a)
            self.row_pruning_method, self.row_pruning_enabled = None, False
        else:  # the *next* layer: drop the matching input columns
            keep = mask.view(-1)
            if dim_reduction:
                self.weight = nn.Parameter(self.weight.data[:, keep])
                self.in_features = int(keep.sum())
            else:
                self.weight.data = self.weight.data * keep.view(1, -1)
            mask = None
        return mask

    def fix_head_pruning_helper(self, mask=None, num_heads=None, dim_reduction=False):
        num_heads = num_heads or self.num_heads
        if mask is None:
            mask = self.get_mask("head").bool().view(-1)
            cols = mask.repeat_interleave(self.in_features // num_heads)
            if dim_reduction:
                self.weight = nn.Parameter(self.weight.data[:, cols])
                self.in_features = int(cols.sum())
            else:
                self.weight.data = self.weight.data * cols.view(1, -1)
            del self.head_pruning_scores
            self.head_pruning_method, self.head_pruning_enabled = None, False
        else:  # qkv projection feeding the pruned heads: drop rows of each of q, k, v
            rows = mask.repeat_interleave(self.out_features // 3 // num_heads).repeat(3)
            if dim_reduction:
                self.weight = nn.Parameter(self.weight.data[rows])
                if self.bias is not None:
                    self.bias = nn.Parameter(self.bias.data[rows])
                self.out_features = int(rows.sum())
            else:
                self.weight.data = self.weight.data * rows.view(-1, 1)
            mask = None
        return mask

    def forward(self, input, skip_bias_add=False):
        weight, bias = self.weight, self.bias
        if self.weight_quantization_enabled_in_forward and self.weight_quantization_enabled:
            weight = self.weight_quantizer(weight, self.weight.target_bits, None, None, self.weight_quantize_num_groups)
        if self.sparse_pruning_enabled and self.sparse_pruning_method:
            weight = weight * self.get_mask("sparse")
        if self.row_pruning_enabled and self.row_pruning_method:
            m = self.get_mask("row")
            weight = weight * m.view(-1, 1)
            if bias is not None:
                bias = bias * m.view(-1)
        if self.head_pruning_enabled and self.head_pruning_method:
            m = self.get_mask("head")
            weight = weight * m.repeat_interleave(self.in_features // self.num_heads, dim=1)
        if self.activation_quantization_enabled:
            input = self._quant_input(input)
        if skip_bias_add:
            return F.linear(input, weight, None), self.bias
        return F.linear(input, weight, bias)


class Conv2dLayer_Compress(nn.Conv2d, _CompressMixin):

    def __init__(self, *kargs):
        super().__init__(*kargs)
        self._init_compress_state()

    def enable_channel_pruning(self, ratio, method):
        self.channel_pruning_ratio, self.channel_pruning_method = ratio, method
        if method == "l1":
            norms = self.weight.detach().abs().sum(dim=(1, 2, 3))
            k = int(norms.numel() * (1 - ratio))
            thr = norms.kthvalue(k).values if k > 0 else norms.min() - 1
            self.register_buffer("channel_pruning_mask", (norms > thr).view(-1, 1, 1, 1))
        elif method == "topk":
            self.channel_mask_scores = nn.Parameter(torch.empty(self.weight.shape[0], 1, 1, 1))
            nn.init.kaiming_uniform_(self.channel_mask_scores, a=math.sqrt(5))
        else:
            raise NotImplementedError

    def fix_channel_pruning_helper(self, mask=None, dim_reduction=False):
        if mask is None:
            mask = self.get_mask("channel").bool().view(-1)
            if dim_reduction:
                self.weight = nn.Parameter(self.weight.data[mask])
                if self.bias is not None:
                    self.bias = nn.Parameter(self.bias.data[mask])
                self.out_channels = int(mask.sum())
            else:
                self.weight.data = self.weight.data * mask.view(-1, 1, 1, 1)
                if self.bias is not None:
                    self.bias.data = self.bias.data * mask
            for a in ("channel_pruning_mask", "channel_mask_scores"):
                if hasattr(self, a):
                    delattr(self, a)
            self.channel_pruning_method, self.channel_pruning_enabled = None, False
        else:
            if dim_reduction:
                self.weight = nn.Parameter(self.weight.data[:, mask])
                self.in_channels = int(mask.sum())
            else:
                self.weight.data = self.weight.data * mask.view(1, -1, 1, 1)
            mask = None
        return mask

    def forward(self, input):
        weight, bias = self.weight, self.bias
        if self.weight_quantization_enabled_in_forward and self.weight_quantization_enabled:
            weight = self.weight_quantizer(weight, self.weight.target_bits, None, None, self.weight_quantize_num_groups)
        if self.sparse_pruning_enabled and self.sparse_pruning_method:
            weight = weight * self.get_mask("sparse")
        if self.channel_pruning_enabled and self.channel_pruning_method:
            m = self.get_mask("channel")
            weight = weight * m
            if bias is not None:
                bias = bias * m.view(-1)
        if self.activation_quantization_enabled:
            input = self._quant_input(input)
        return F.conv2d(input, weight, bias, self.stride, self.padding, self.dilation, self.groups)


class BNLayer_Compress(nn.BatchNorm2d):

    def fix_channel_pruning_helper(self, mask, dim_reduction=True):
        for name in ("weight", "bias"):
            setattr(self, name, nn.Parameter(getattr(self, name).data[mask.view(-1)]))
        self.running_mean = self.running_mean[mask.view(-1)]
        self.running_var = self.running_var[mask.view(-1)]
        self.num_features = int(mask.sum())


class Embedding_Compress(nn.Embedding, _CompressMixin):

    def __init__(self, *kargs):
        super().__init__(*kargs)
        self._init_compress_state()
        self.weight_quantize_num_groups = self.weight.size(0)

    def enable_weight_quantization(self, start_bits, target_bits, quantization_period, weight_quantization_enabled_in_forward,
                                   quantization_type, num_groups):
        super().enable_weight_quantization(start_bits, target_bits, quantization_period,
                                           weight_quantization_enabled_in_forward, quantization_type, self.weight.size(0))

    def forward(self, input):
        weight = self.weight
        if self.weight_quantization_enabled_in_forward and self.weight_quantization_enabled:
            weight = self.weight_quantizer(weight, self.weight.target_bits, None, None, self.weight_quantize_num_groups)
        return F.embedding(input, weight, self.padding_idx, self.max_norm, self.norm_type, self.scale_grad_by_freq, self.sparse)


class ColumnParallelLinear_Compress(LinearLayer_Compress):
    """Megatron column-parallel linear with compression (output features sharded over ``mpu``'s TP group)."""

    def __init__(self, mpu, input_size, output_size, bias=True, gather_output=True, skip_bias_add=False):
        self.mpu = mpu
        world = mpu.get_model_parallel_world_size()
        assert output_size % world == 0
        super().__init__(input_size, output_size // world, bias=bias)
        self.gather_output, self.skip_bias_add = gather_output, skip_bias_add

    def forward(self, input_):
        from deepspeed_b200.module_inject.layers import ColumnParallel, _gather_dim
        x = ColumnParallel.apply(self.mpu.get_model_parallel_group(), input_)
        if self.skip_bias_add:
            out, bias = super().forward(x, True)
        else:
            out, bias = super().forward(x), None
        if self.gather_output:
            out = _gather_dim(out, self.mpu.get_model_parallel_group(), out.dim() - 1)
        return (out, bias) if self.skip_bias_add else out


class RowParallelLinear_Compress(LinearLayer_Compress):

    def __init__(self, mpu, input_size, output_size, bias=True, input_is_parallel=False, skip_bias_add=False):
        self.mpu = mpu
        world = mpu.get_model_parallel_world_size()
        assert input_size % world == 0
        super().__init__(input_size // world, output_size, bias=bias)
        self.input_is_parallel, self.skip_bias_add = input_is_parallel, skip_bias_add

    def forward(self, input_):
        from deepspeed_b200.module_inject.layers import RowParallel
        if not self.input_is_parallel:
            w, r = self.mpu.get_model_parallel_world_size(), self.mpu.get_model_parallel_rank()
            input_ = input_.chunk(w, dim=-1)[r].contiguous()
        out, bias = super().forward(input_, True)
        out = RowParallel.apply(self.mpu.get_model_parallel_group(), out)
        if self.skip_bias_add:
            return out, bias
        return out + bias if bias is not None else out


# --- Megatron-style model-parallel region operators (reference ``compression/basic_layer.py:620-765``) ----------------
# ``mpu``-less functional forms: they act on the framework's tensor-parallel group (``utils.groups``), which is what the
# ``*_Compress`` parallel layers above use when constructed without an explicit mpu.
def _tp_group():
    from deepspeed_b200.utils import groups
    return groups.get_tensor_model_parallel_group()


def _tp_world(group):
    from deepspeed_b200 import comm as dist
    return dist.get_world_size(group=group) if dist.is_initialized() else 1


def split_tensor_along_last_dim(tensor, num_partitions, contiguous_split_chunks=False):
    assert tensor.size(-1) % num_partitions == 0, f"{tensor.size(-1)} is not divisible by {num_partitions}"
    parts = tensor.split(tensor.size(-1) // num_partitions, dim=-1)
    return tuple(p.contiguous() for p in parts) if contiguous_split_chunks else parts


def _reduce(input_, group=None):
    from deepspeed_b200 import comm as dist
    group = group if group is not None else _tp_group()
    if _tp_world(group) > 1:
        dist.all_reduce(input_, group=group)
    return input_


def _split(input_, group=None):
    from deepspeed_b200 import comm as dist
    group = group if group is not None else _tp_group()
    w = _tp_world(group)
    return input_ if w == 1 else split_tensor_along_last_dim(input_, w)[dist.get_rank(group=group)].contiguous()


def _gather(input_, group=None):
    from deepspeed_b200 import comm as dist
    group = group if group is not None else _tp_group()
    w = _tp_world(group)
    if w == 1:
        return input_
    parts = [torch.empty_like(input_) for _ in range(w)]
    dist.all_gather(parts, input_.contiguous(), group=group)
    return torch.cat(parts, dim=-1)


def _region(name, fwd, bwd):
    """Autograd function whose forward is ``fwd`` and whose backward applies ``bwd`` to the incoming gradient."""

    def forward(ctx, x):
        return fwd(x)

    def backward(ctx, g):
        return bwd(g)

    return type(name, (torch.autograd.Function, ), {"forward": staticmethod(forward), "backward": staticmethod(backward)})


_identity = lambda t: t
_CopyToModelParallelRegion = _region("_CopyToModelParallelRegion", _identity, lambda g: _reduce(g.clone()))
_ReduceFromModelParallelRegion = _region("_ReduceFromModelParallelRegion", lambda x: _reduce(x.clone()), _identity)
_ScatterToModelParallelRegion = _region("_ScatterToModelParallelRegion", _split, _gather)
_GatherFromModelParallelRegion = _region("_GatherFromModelParallelRegion", _gather, _split)


def copy_to_model_parallel_region(input_):
    return _CopyToModelParallelRegion.apply(input_)


def reduce_from_model_parallel_region(input_):
    return _ReduceFromModelParallelRegion.apply(input_)


def scatter_to_model_parallel_region(input_):
    return _ScatterToModelParallelRegion.apply(input_)


def gather_from_model_parallel_region(input_):
    return _GatherFromModelParallelRegion.apply(input_)
