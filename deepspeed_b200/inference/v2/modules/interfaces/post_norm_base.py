"""``DSPostNormBase`` interface + ``DSPostNormRegistry`` (reference ``modules/interfaces/post_norm_base.py``)."""
from typing import Any, Dict, Type

import torch

from ..configs.norm_config import DSNormConfig
from ..ds_module import DSModuleBase
from ..module_registry import DSModuleRegistryBase


class DSPostNormBase(DSModuleBase):

    @staticmethod
    def config_class() -> Type[DSNormConfig]:
        return DSNormConfig

    def __init__(self, config: DSNormConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)

    def transform_param(self, param: torch.Tensor):
        return param

    def forward(self, residual: torch.Tensor, hidden_in: torch.Tensor, gamma, beta=None) -> torch.Tensor:
        """``norm(residual + hidden_in)`` - the result is both the next residual and the next hidden state."""
        raise NotImplementedError


class DSPostNormRegistry(DSModuleRegistryBase):

    @classmethod
    def associated_class(cls) -> Type[DSModuleBase]:
        return DSPostNormBase
