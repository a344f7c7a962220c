"""``DSPreNormBase`` interface + ``DSPreNormRegistry`` (reference ``modules/interfaces/pre_norm_base.py``)."""
from typing import Any, Dict, Type

import torch

from ..configs.norm_config import DSNormConfig
from ..ds_module import DSModuleBase
from ..module_registry import DSModuleRegistryBase


class DSPreNormBase(DSModuleBase):

    @staticmethod
    def config_class() -> Type[DSNormConfig]:
        return DSNormConfig

    def __init__(self, config: DSNormConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)

    def transform_param(self, param: torch.Tensor):
        return param

    def forward(self, residual: torch.Tensor, hidden_in, gamma, beta=None):
        """``residual += hidden_in`` (skipped when ``hidden_in`` is None); returns (residual, norm(residual))."""
        raise NotImplementedError


class DSPreNormRegistry(DSModuleRegistryBase):

    @classmethod
    def associated_class(cls) -> Type[DSModuleBase]:
        return DSPreNormBase
