"""Softmax + top-k expert selection per token, with the per-expert counts and each copy's slot inside its expert.

Reference ``inference/v2/kernels/ragged_ops/top_k_gating/top_k_gating.py``."""
import torch

from deepspeed_b200.ops.kernels import moe_ops as M

from ...ds_kernel import DSKernelBase, check_dtype


class RaggedTopKGating(DSKernelBase):
    supported_logit_dtypes = [torch.float16, torch.bfloat16, torch.float32]

    def __init__(self, logit_dtype) -> None:
        check_dtype(logit_dtype, "RaggedTopKGating")

    def __call__(self, expert_counts, scores, assignments, offsets, logits, batch=None):
        """Fills ``scores`` / ``assignments`` / ``offsets`` [T, k] and ``expert_counts`` [E]."""
        T, k = assignments.shape
        ids, w, _ = M.top_k_gating(logits, k, normalize=False)
        positions, counts, _ = M.route(ids, expert_counts.numel())
        scores.copy_(w.view(T, k))
        assignments.copy_(ids.view(T, k))
        offsets.copy_(positions.view(T, k))
        expert_counts.copy_(counts)
        return expert_counts, scores, assignments, offsets
