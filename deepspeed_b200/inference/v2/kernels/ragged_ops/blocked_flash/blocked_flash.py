"""Attention of a ragged batch over the blocked (paged) KV cache: every query token attends to keys ``[0, pos]`` of its sequence; GQA-shared split-KV kernel at decode sizes.

Reference ``inference/v2/kernels/ragged_ops/blocked_flash/blocked_flash.py``."""
import torch

from deepspeed_b200.ops.kernels import ragged_ops as R

from ...ds_kernel import DSKernelBase, check_dtype


class BlockedFlashAttn(DSKernelBase):
    supported_dtypes = [torch.float16, torch.bfloat16]

    def __init__(self, head_size: int, dtype) -> None:
        check_dtype(dtype, "BlockedFlashAttn")
        if head_size % 16 != 0:
            raise ValueError("Head size must be divisible by 16")
        self.head_size = head_size

    def __call__(self, out, qkv, kv_cache, seq_of, pos_of, block_table, n_q_heads, n_kv_heads, block_size, softmax_scale=None):
        """``qkv`` [T, (hq + 2 hkv) d] (K/V already appended to ``kv_cache`` by the rotary kernel)."""
        out.copy_(R.paged_attention(qkv, kv_cache, seq_of, pos_of, block_table, n_q_heads, n_kv_heads, self.head_size, block_size,
                                    softmax_scale))
        return out
