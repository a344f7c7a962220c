"""Append k, v to the blocked KV cache without rotation (learned / ALiBi positions).

Reference ``inference/v2/kernels/ragged_ops/linear_blocked_kv_rotary/linear_blocked_kv_copy.py``."""
import torch

from deepspeed_b200.ops.kernels import ragged_ops as R

from ...ds_kernel import DSKernelBase, check_dtype


class LinearBlockedKVCopy(DSKernelBase):

    def __init__(self, head_size: int, n_q_heads: int, n_kv_heads: int, dtype) -> None:
        check_dtype(dtype, "LinearBlockedKVCopy")
        self.head_size, self.hq, self.hkv = head_size, n_q_heads, n_kv_heads

    def __call__(self, kv_cache, qkv, seq_of, pos_of, block_table, block_size) -> None:
        R.kv_rotary_append(qkv, kv_cache, None, None, seq_of, pos_of, block_table, self.hq, self.hkv, self.head_size, 0, block_size)
