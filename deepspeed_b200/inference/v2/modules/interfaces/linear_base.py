"""``DSLinearBase`` interface + ``DSLinearRegistry`` (reference ``modules/interfaces/linear_base.py``)."""
from typing import Any, Dict, Type

import torch

from ..configs.linear_config import DSLinearConfig
from ..ds_module import DSModuleBase
from ..module_registry import DSModuleRegistryBase


class DSLinearBase(DSModuleBase):

    @staticmethod
    def config_class() -> Type[DSLinearConfig]:
        return DSLinearConfig

    def __init__(self, config: DSLinearConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)

    def transform_param(self, param: torch.Tensor):
        """Put a checkpoint weight / bias into the layout (or quantised form) this implementation multiplies with."""
        return param

    def forward(self, hidden_states: torch.Tensor, w, b=None) -> torch.Tensor:
        raise NotImplementedError

    @property
    def output(self) -> torch.Tensor:
        raise NotImplementedError


class DSLinearRegistry(DSModuleRegistryBase):

    @classmethod
    def associated_class(cls) -> Type[DSModuleBase]:
        return DSLinearBase
