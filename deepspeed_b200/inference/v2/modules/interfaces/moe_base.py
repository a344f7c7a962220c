"""``DSMoEBase`` interface + ``DSMoERegistry`` (reference ``modules/interfaces/moe_base.py``)."""
from typing import Any, Dict, Type

import torch

from ..configs.moe_config import DSMoEConfig
from ..ds_module import DSModuleBase
from ..module_registry import DSModuleRegistryBase


class DSMoEBase(DSModuleBase):

    @staticmethod
    def config_class() -> Type[DSMoEConfig]:
        return DSMoEConfig

    def __init__(self, config: DSMoEConfig, implementation_config: Dict[str, Any] = None) -> None:
        super().__init__(config, implementation_config)

    def transform_gate_param(self, param: torch.Tensor):
        return param

    def transform_moe_mlp_1_param(self, param: torch.Tensor):
        return param

    def transform_moe_mlp_2_param(self, param: torch.Tensor):
        return param

    def forward(self, hidden_states, gate_w, mlp_1_w, mlp_2_w, mlp_1_b=None, mlp_2_b=None) -> torch.Tensor:
        raise NotImplementedError

    @property
    def output(self) -> torch.Tensor:
        raise NotImplementedError


class DSMoERegistry(DSModuleRegistryBase):

    @classmethod
    def associated_class(cls) -> Type[DSModuleBase]:
        return DSMoEBase
