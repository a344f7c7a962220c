"""``TiledLinear``: split a huge linear into an ``in_splits x out_splits`` grid of small linears so ZeRO-3 only
ever materialises one tile's weights at a time (reference ``runtime/zero/tiling.py:32``)."""
import torch
from torch import nn

from deepspeed_b200.runtime.utils import partition_uniform


def split_tensor_along_last_dim(tensor, partitions, contiguous_split_chunks=False):
    sizes = [partitions[i + 1] - partitions[i] for i in range(len(partitions) - 1)]
    parts = torch.split(tensor, sizes, dim=tensor.dim() - 1)
    return tuple(p.contiguous() for p in parts) if contiguous_split_chunks else parts


class TiledLinear(nn.Module):

    def __init__(self, in_features, out_features, bias=True, in_splits=1, out_splits=1, input_is_already_split=False,
                 combine_out_splits=True, linear_cls=nn.Linear, init_linear=None, **kwargs):
        super().__init__()
        if in_splits < 1 or in_splits > in_features:
            raise RuntimeError("in splits must be in range [1, in_features].")
        if out_splits < 1 or out_splits > out_features:
            raise RuntimeError("out splits must be in range [1, out_features].")
        self.in_features, self.out_features, self.use_bias = in_features, out_features, bias
        self.in_splits, self.out_splits = in_splits, out_splits
        self.input_is_already_split = input_is_already_split
        self.combine_out_splits = combine_out_splits
        self.in_parts = partition_uniform(in_features, in_splits)
        self.out_parts = partition_uniform(out_features, out_splits)
        self.linears = nn.ModuleList()
        for o in range(out_splits):
            row = nn.ModuleList()
            rows = self.out_parts[o + 1] - self.out_parts[o]
            for i in range(in_splits):
                cols = self.in_parts[i + 1] - self.in_parts[i]
                # only the last tile of each row carries the bias so it is added exactly once
                local = linear_cls(cols, rows, bias=bias and i == in_splits - 1, **kwargs)
                row.append(local)
            self.linears.append(row)
        if init_linear is not None:
            self.copy_params_from(init_linear)

    def forward(self, input_):
        if self.in_splits > 1 and not self.input_is_already_split:
            inputs = split_tensor_along_last_dim(input_, self.in_parts)
        elif self.in_splits > 1:
            inputs = input_
            assert len(inputs) == self.in_splits, f"Col splits {self.in_splits} does not match input splits {len(inputs)}"
        else:
            inputs = [input_]
        outputs = []
        for o in range(self.out_splits):
            acc = None
            for i in range(self.in_splits):
                y = self.linears[o][i](inputs[i])
                if isinstance(y, tuple):
                    y = y[0]
                acc = y if acc is None else acc + y
            outputs.append(acc)
        return self._combine(outputs) if self.combine_out_splits else outputs

    def _combine(self, outputs):
        return torch.cat(outputs, dim=-1)

    @torch.no_grad()
    def copy_params_from(self, other):
        assert hasattr(other, "weight") and other.weight.size() == (self.out_features, self.in_features)
        if self.use_bias:
            assert other.bias is not None and other.bias.size() == (self.out_features, )
        else:
            assert other.bias is None
        from deepspeed_b200.runtime.zero.partition_parameters import GatheredParameters
        for o in range(self.out_splits):
            r0, r1 = self.out_parts[o], self.out_parts[o + 1]
            for i in range(self.in_splits):
                c0, c1 = self.in_parts[i], self.in_parts[i + 1]
                local = self.linears[o][i]
                with GatheredParameters(list(local.parameters()), modifier_rank=0):
                    local.weight.copy_(other.weight[r0:r1, c0:c1])
                    if local.bias is not None:
                        local.bias.copy_(other.bias[r0:r1])


class TiledLinearReturnBias(TiledLinear):
    """For Megatron-style linears whose forward returns ``(output, bias)``."""

    def forward(self, input_):
        if self.in_splits > 1 and not self.input_is_already_split:
            inputs = split_tensor_along_last_dim(input_, self.in_parts)
        elif self.in_splits > 1:
            inputs = input_
        else:
            inputs = [input_]
        outs, biases = [], []
        for o in range(self.out_splits):
            acc, b = None, None
            for i in range(self.in_splits):
                y = self.linears[o][i](inputs[i])
                yb = None
                if isinstance(y, tuple):
                    y, yb = y
                acc = y if acc is None else acc + y
                b = yb if yb is not None else b
            outs.append(acc)
            biases.append(b)
        if self.combine_out_splits:
            bias = torch.cat(biases, dim=-1) if all(b is not None for b in biases) else None
            return torch.cat(outs, dim=-1), bias
        return outs, biases
