"""Token gather / drop across the tensor-parallel group around an MoE block (reference ``moe/mappings.py``).

With TP > 1 every TP rank holds the same tokens; ``drop_tokens`` keeps this rank's 1/tp slice before the expert
all-to-all (so experts do not see duplicates) and ``gather_tokens`` restores the full set afterwards.  Each is the
other's backward.
"""
import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.utils import groups
from deepspeed_b200.utils.bwc import (bwc_tensor_model_parallel_group, bwc_tensor_model_parallel_rank,
                                      bwc_tensor_model_parallel_world_size)


def _mpu():
    return groups._mpu


def _tp():
    mpu = _mpu()
    return bwc_tensor_model_parallel_world_size(mpu), bwc_tensor_model_parallel_rank(mpu), bwc_tensor_model_parallel_group(mpu)


def _gather_tokens(input_, dim=0):
    world, _, group = _tp()
    if world == 1:
        return input_
    x = input_.movedim(dim, 0).contiguous()
    out = torch.empty((world * x.shape[0], *x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x, group=group)
    return out.movedim(0, dim).contiguous()


def _drop_tokens(input_, dim=0):
    world, rank, _ = _tp()
    if world == 1:
        return input_
    n = input_.shape[dim]
    assert n % world == 0, f"input dimension {dim} ({n}) is not divisible by tensor parallel world size ({world})"
    return input_.narrow(dim, rank * (n // world), n // world)


class _GatherTokens(torch.autograd.Function):

    @staticmethod
    def forward(ctx, input_, dim):
        ctx.dim = dim
        return _gather_tokens(input_, dim)

    @staticmethod
    def backward(ctx, grad_output):
        return _drop_tokens(grad_output, ctx.dim), None


class _DropTokens(torch.autograd.Function):

    @staticmethod
    def forward(ctx, input_, dim):
        ctx.dim = dim
        return _drop_tokens(input_, dim)

    @staticmethod
    def backward(ctx, grad_output):
        return _gather_tokens(grad_output, ctx.dim), None


def gather_tokens(input_, dim=0):
    if _tp()[0] == 1:
        return input_
    return _GatherTokens.apply(input_, dim)


def drop_tokens(input_, dim=0):
    if _tp()[0] == 1:
        return input_
    return _DropTokens.apply(input_, dim)
