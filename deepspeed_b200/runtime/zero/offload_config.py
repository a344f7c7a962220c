"""ZeRO offload config blocks (reference: ``runtime/zero/offload_config.py:21,51``)."""
from enum import Enum
from pathlib import Path
from typing import Optional

from pydantic import Field, model_validator

from deepspeed_b200.runtime.config_utils import DeepSpeedConfigModel, pp_int


class OffloadDeviceEnum(str, Enum):
    none = "none"
    cpu = "cpu"
    nvme = "nvme"


class OffloadStateTypeEnum(str, Enum):
    """States addressable by ``engine.offload_states`` (reference: offload_states.py)."""
    optim_states = "optim_states"
    hp_params = "hp_params"
    lp_params = "lp_params"
    lp_grads = "lp_grads"
    contiguous_grad_buffer = "contiguous_grad_buffer"


class DeepSpeedZeroOffloadParamConfig(DeepSpeedConfigModel):
    device: OffloadDeviceEnum = "none"
    nvme_path: Optional[Path] = None
    buffer_count: int = Field(5, ge=0)
    buffer_size: int = Field(pp_int(1e8), ge=0)
    max_in_cpu: int = Field(pp_int(1e9), ge=0)
    pin_memory: bool = False


class DeepSpeedZeroOffloadOptimizerConfig(DeepSpeedConfigModel):
    device: OffloadDeviceEnum = "none"
    nvme_path: Optional[Path] = None
    buffer_count: int = Field(4, ge=0)
    pin_memory: bool = False
    pipeline_read: bool = False
    pipeline_write: bool = False
    fast_init: bool = False
    ratio: float = Field(1.0, ge=0.0, le=1.0)  # Twin-Flow: fraction of optimizer state on the host
    b200_swap_window: int = Field(pp_int(1 << 26), ge=1)  # NVMe tier: elements per pinned streaming window
    b200_swap_master: bool = True  # NVMe tier: the fp32 master weights live in a swap file too (reference behaviour)

    @model_validator(mode="after")
    def set_pipeline(self):
        self.__dict__["pipeline"] = self.pipeline_read or self.pipeline_write
        return self

    @property
    def pipeline(self):
        return self.pipeline_read or self.pipeline_write
