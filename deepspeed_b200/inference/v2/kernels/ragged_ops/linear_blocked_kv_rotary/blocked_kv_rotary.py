"""Rotate q and k by tables computed from ``theta_base`` and append k, v to the blocked KV cache, in one pass over the packed QKV.

Reference ``inference/v2/kernels/ragged_ops/linear_blocked_kv_rotary/blocked_kv_rotary.py``."""
import torch

from deepspeed_b200.ops.kernels import ragged_ops as R
from deepspeed_b200.ops.kernels.transformer_ops import RotaryTable

from ...ds_kernel import DSKernelBase, check_dtype


class BlockedRotaryEmbeddings(DSKernelBase):
    supported_head_sizes = [64, 80, 96, 128]

    def __init__(self, head_size: int, n_q_heads: int, n_kv_heads: int, dtype, rotary_dim: int, theta_base: float,
                 max_positions: int = 8192) -> None:
        check_dtype(dtype, "BlockedRotaryEmbeddings")
        if n_q_heads % n_kv_heads != 0:
            raise ValueError("n_q_heads must be a multiple of n_kv_heads")
        self.head_size, self.hq, self.hkv, self.rot = head_size, n_q_heads, n_kv_heads, rotary_dim
        self.table = RotaryTable(rotary_dim, max_positions, base=theta_base)

    def __call__(self, kv_cache, qkv, seq_of, pos_of, block_table, block_size) -> None:
        if self.table.cos.device != qkv.device:
            self.table.to(qkv.device)
        R.kv_rotary_append(qkv, kv_cache, self.table.cos, self.table.sin, seq_of, pos_of, block_table, self.hq, self.hkv,
                           self.head_size, self.rot, block_size)
