"""Token (+ learned position) embedding lookup for a ragged batch.

Reference ``inference/v2/kernels/ragged_ops/embed/embed.py``."""
import torch

from deepspeed_b200.ops.kernels import ragged_ops as R

from ...ds_kernel import DSKernelBase, check_dtype


class RaggedEmbeddingKernel(DSKernelBase):
    supported_token_dtypes = [torch.int32, torch.int64]

    def __init__(self, embed_dtype, token_dtype, embed_dim: int) -> None:
        check_dtype(embed_dtype, "RaggedEmbeddingKernel")
        if token_dtype not in self.supported_token_dtypes:
            raise ValueError(f"Unsupported token dtype {token_dtype}")
        if embed_dim * torch.empty(0, dtype=embed_dtype).element_size() % 16 != 0:
            raise ValueError("embedding dim must be a multiple of 16 bytes")

    def __call__(self, embedded_tokens, token_ids, embedding_weight, position_ids=None, position_embed_weight=None,
                 position_embed_offset=0) -> torch.Tensor:
        embedded_tokens.copy_(R.ragged_embed(token_ids, embedding_weight, position_ids, position_embed_weight, position_embed_offset))
        return embedded_tokens
