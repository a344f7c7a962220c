"""Blocked-KV dense attention (reference ``modules/implementations/attention/dense_blocked_attention.py``):
rotary (or plain) KV append + paged attention over the ragged batch."""
from typing import Any, Dict

import torch

from ....inference_utils import DtypeEnum
from ....kernels.ragged_ops import (BlockedFlashAttn, BlockedRotaryEmbeddings, BlockedTrainedRotaryEmbeddings,
                                    LinearBlockedKVCopy)
from ...configs import DSSelfAttentionConfig, MaskingType, PositionalEmbeddingType
from ...interfaces import DSSelfAttentionBase, DSSelfAttentionRegistry


@DSSelfAttentionRegistry.register_module
class DSDenseBlockedAttention(DSSelfAttentionBase):

    @staticmethod
    def name() -> str:
        return "dense_blocked_attention"

    @staticmethod
    def supports_config(config: DSSelfAttentionConfig) -> bool:
        if config.input_dtype != config.output_dtype or config.n_heads_q % config.n_heads_kv != 0:
            return False
        if MaskingType(config.masking_type) != MaskingType.causal:
            return False
        return PositionalEmbeddingType(config.positional_embedding_type) in (PositionalEmbeddingType.none,
                                                                             PositionalEmbeddingType.rotate_half)

    def __init__(self, config: DSSelfAttentionConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)
        c = config
        dt = DtypeEnum(c.input_dtype).value
        self._trained = False
        if PositionalEmbeddingType(c.positional_embedding_type) == PositionalEmbeddingType.none:
            self.kv_op = LinearBlockedKVCopy(c.head_size, c.n_heads_q, c.n_heads_kv, dt)
        else:
            rc = c.positional_embedding_config
            if rc is not None and rc.use_trained_freqs:
                self.kv_op = BlockedTrainedRotaryEmbeddings(c.head_size, c.n_heads_q, c.n_heads_kv, dt)
                self._trained = True
            else:
                self.kv_op = BlockedRotaryEmbeddings(c.head_size, c.n_heads_q, c.n_heads_kv, dt,
                                                     (rc.rotate_dim if rc is not None and rc.rotate_dim else c.head_size),
                                                     rc.theta_base if rc is not None else 10000.0,
                                                     max_positions=(implementation_config or {}).get("max_positions", 8192))
        self.attn = BlockedFlashAttn(c.head_size, dt)
        self._block = int((implementation_config or {}).get("kv_block_size", 128))

    @property
    def kv_block_size(self) -> int:
        return self._block

    def forward(self, q_k_v, kv_cache, batch, inv_freqs=None) -> torch.Tensor:
        c = self._config
        seq_of, pos_of, bt = batch.seq_of(), batch.pos_of(), batch.block_table()
        if self._trained:
            cos, sin = inv_freqs
            self.kv_op(kv_cache, q_k_v, seq_of, pos_of, bt, self._block, cos, sin)
        else:
            self.kv_op(kv_cache, q_k_v, seq_of, pos_of, bt, self._block)
        out = torch.empty(q_k_v.shape[0], c.n_heads_q * c.head_size, dtype=q_k_v.dtype, device=q_k_v.device)
        return self.attn(out, q_k_v, kv_cache, seq_of, pos_of, bt, c.n_heads_q, c.n_heads_kv, self._block,
                         c.scale_factor if c.scale_factor != 1.0 else None)
