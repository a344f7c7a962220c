"""As ``BlockedRotaryEmbeddings`` but with caller-provided (trained / scaled) cos-sin tables.

Reference ``inference/v2/kernels/ragged_ops/linear_blocked_kv_rotary/blocked_trained_kv_rotary.py``."""
import torch

from deepspeed_b200.ops.kernels import ragged_ops as R

from ...ds_kernel import DSKernelBase, check_dtype


class BlockedTrainedRotaryEmbeddings(DSKernelBase):

    def __init__(self, head_size: int, n_q_heads: int, n_kv_heads: int, dtype) -> None:
        check_dtype(dtype, "BlockedTrainedRotaryEmbeddings")
        self.head_size, self.hq, self.hkv = head_size, n_q_heads, n_kv_heads

    def __call__(self, kv_cache, qkv, seq_of, pos_of, block_table, block_size, cos, sin) -> None:
        R.kv_rotary_append(qkv, kv_cache, cos, sin, seq_of, pos_of, block_table, self.hq, self.hkv, self.head_size,
                           2 * cos.shape[-1], block_size)
