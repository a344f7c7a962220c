"""``DSSelfAttentionBase`` interface + ``DSSelfAttentionRegistry`` (reference ``modules/interfaces/attention_base.py``)."""
from typing import Any, Dict, Type

import torch

from ..configs.attention_configs import DSSelfAttentionConfig
from ..ds_module import DSModuleBase
from ..module_registry import DSModuleRegistryBase


class DSSelfAttentionBase(DSModuleBase):

    @staticmethod
    def config_class() -> Type[DSSelfAttentionConfig]:
        return DSSelfAttentionConfig

    def __init__(self, config: DSSelfAttentionConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)

    @property
    def kv_block_size(self) -> int:
        """Tokens per KV-cache block this implementation wants."""
        raise NotImplementedError

    def forward(self, q_k_v: torch.Tensor, kv_cache: torch.Tensor, batch, inv_freqs=None) -> torch.Tensor:
        """``q_k_v`` [tokens, (hq + 2 hkv) * d] packed; ``batch`` carries seq_of / pos_of / block_table."""
        raise NotImplementedError


class DSSelfAttentionRegistry(DSModuleRegistryBase):

    @classmethod
    def associated_class(cls) -> Type[DSModuleBase]:
        return DSSelfAttentionBase
