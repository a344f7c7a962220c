"""Permute token activations into expert-sorted order (one copy per selected expert).

Reference ``inference/v2/kernels/ragged_ops/moe_scatter/moe_scatter.py``."""
import torch

from deepspeed_b200.ops.kernels import moe_ops as M

from ...ds_kernel import DSKernelBase, check_dtype


class MoEScatter(DSKernelBase):

    def __init__(self, dtype, channels: int) -> None:
        check_dtype(dtype, "MoEScatter")
        if channels % 8 != 0:
            raise ValueError("channels must be divisible by 8")

    def __call__(self, moe_input, expert_cumsum, mapped_slots, activations, expert_counts, assignments, offsets):
        """``assignments`` [T, k] expert ids, ``offsets`` [T, k] slot of each copy inside its expert.  Fills ``moe_input``
        (sorted rows), ``expert_cumsum`` (inclusive row ends per expert) and ``mapped_slots`` [T, k]."""
        T, k = assignments.shape
        starts = torch.cumsum(expert_counts, 0) - expert_counts
        off = torch.cat([starts, expert_counts.sum().view(1)]).to(torch.int32)
        rows, slots = M.scatter(activations, assignments.to(torch.int32), offsets.reshape(-1).to(torch.int32), off, k, 0, T * k)
        moe_input[:rows.shape[0]].copy_(rows)
        expert_cumsum.copy_(torch.cumsum(expert_counts, 0))
        mapped_slots.copy_(slots.view(T, k))
        return moe_input, expert_cumsum, mapped_slots
