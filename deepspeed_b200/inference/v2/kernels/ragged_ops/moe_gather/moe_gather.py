"""Un-permute expert outputs back to token order and combine the top-k copies with their gate weights.

Reference ``inference/v2/kernels/ragged_ops/moe_gather/moe_gather.py``."""
import torch

from deepspeed_b200.ops.kernels import moe_ops as M

from ...ds_kernel import DSKernelBase, check_dtype


class MoEGather(DSKernelBase):

    def __init__(self, dtype, channels: int, normalize_scores: bool = False) -> None:
        check_dtype(dtype, "MoEGather")
        if channels % 8 != 0:
            raise ValueError("channels must be divisible by 8")
        self.normalize_scores = normalize_scores

    def __call__(self, layer_output, moe_output, scores, mapped_slots, expert_counts=None) -> torch.Tensor:
        T, k = scores.shape
        w = scores.float()
        if self.normalize_scores:
            w = w / w.sum(dim=-1, keepdim=True)
        layer_output.copy_(M.gather(moe_output, w, mapped_slots.reshape(-1), T, k))
        return layer_output
