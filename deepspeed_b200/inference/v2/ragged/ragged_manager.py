"""Sequence / KV state manager (reference ``ragged/ragged_manager.py:19 DSStateManager``)."""
from typing import Dict, Optional, Tuple

import torch

from deepspeed_b200.utils.logging import logger
from ..config_v2 import DSStateManagerConfig, KVCacheConfig
from .kv_cache import BlockedKVCache
from .sequence_descriptor import DSSequenceDescriptor


class DSStateManager:

    def __init__(self, config: DSStateManagerConfig, kv_configs: Tuple[KVCacheConfig, ...], base_mp_group=None,
                 device=None):
        self._config = config
        self._kv_configs = kv_configs
        self._seqs: Dict[int, DSSequenceDescriptor] = {}
        self._kv_cache = BlockedKVCache(kv_configs, config.memory_config, mp_group=base_mp_group, offload=config.offload,
                                        device=device)

    def get_cache(self, cache_id: int, cache_group: int = 0) -> torch.Tensor:
        return self._kv_cache.get_cache(cache_id, cache_group)

    def query(self, uid: Optional[int] = None):
        if uid is None:
            return self._seqs
        return self._seqs.get(uid)

    def get_sequence(self, uid: int) -> Optional[DSSequenceDescriptor]:
        return self._seqs.get(uid)

    def get_or_create_sequence(self, uid: int) -> DSSequenceDescriptor:
        seq = self._seqs.get(uid)
        if seq is not None:
            return seq
        if len(self._seqs) >= self._config.max_tracked_sequences:
            raise RuntimeError(f"Too many tracked sequences ({len(self._seqs)})")
        ids = tuple(torch.zeros(c.max_blocks_per_allocation_group, dtype=torch.int32) for c in self._kv_configs)
        seq = DSSequenceDescriptor(uid, ids, max_context=self._config.max_context)
        self._seqs[uid] = seq
        return seq

    def flush_sequence(self, uid: int) -> None:
        seq = self._seqs.pop(uid, None)
        if seq is None:
            logger.warning(f"Attempting to flush sequence {uid} which does not exist.")
            return
        for g in range(self.n_kv_cache_groups):
            self._kv_cache.free(seq.all_block_ids(g), cache_group=g)

    def offload_sequence(self, uid: int) -> None:
        """Move a paused sequence's KV blocks to pinned host memory and release its device blocks."""
        seq = self._seqs[uid]
        blocks = seq.all_block_ids(0).clone()
        seq.host_kv = self._kv_cache.offload(blocks.tolist())
        self._kv_cache.free(blocks)
        seq._blocks_per[0] = 0

    def restore_sequence(self, uid: int) -> None:
        seq = self._seqs[uid]
        if seq.host_kv is None:
            return
        n = seq.host_kv.shape[1]
        blocks = self._kv_cache.reserve(n)
        self._kv_cache.restore(blocks.tolist(), seq.host_kv)
        seq.extend_kv_cache(blocks)
        seq.host_kv = None

    def allocate_blocks(self, n_blocks: int, cache_group: int = 0) -> torch.Tensor:
        return self._kv_cache.reserve(n_blocks, cache_group)

    @property
    def tracked_sequences(self):
        return self._seqs

    @property
    def n_tracked_sequences(self) -> int:
        return len(self._seqs)

    @property
    def kv_block_size(self) -> int:
        return self._kv_configs[0].block_size

    @property
    def n_kv_cache_groups(self) -> int:
        return self._kv_cache.num_caches

    def free_block_count(self, cache_group: int = 0) -> int:
        return self._kv_cache.free_block_count(cache_group)

    @property
    def free_blocks(self) -> torch.Tensor:
        return self._kv_cache.free_blocks
