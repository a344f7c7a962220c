"""Inference (v1 API) config.  Reference: ``inference/config.py`` ``DeepSpeedInferenceConfig``."""
from enum import Enum
from typing import Dict, Optional, Union

import torch
from pydantic import Field, field_validator

from deepspeed_b200.runtime.config_utils import DeepSpeedConfigModel
from deepspeed_b200.runtime.zero.config import DeepSpeedZeroConfig


class DtypeEnum(Enum):
    fp16 = (torch.float16, "torch.float16", "fp16", "float16", "half")
    fp32 = (torch.float32, "torch.float32", "fp32", "float32", "float")
    bf16 = (torch.bfloat16, "torch.bfloat16", "bf16", "bfloat16", "bfloat")
    int8 = (torch.int8, "torch.int8", "int8")

    @classmethod
    def from_any(cls, v):
        if isinstance(v, cls):
            return v
        for m in cls:
            if v in m.value or v is m.value[0]:
                return m
        raise ValueError(f"unknown dtype {v}")

    @property
    def torch(self):
        return self.value[0]


class MoETypeEnum(str, Enum):
    residual = "residual"
    standard = "standard"


class DeepSpeedTPConfig(DeepSpeedConfigModel):
    enabled: bool = True
    tp_size: int = 1
    tp_grain_size: int = 64
    mpu: object = None
    tp_group: object = None


class DeepSpeedMoEConfig(DeepSpeedConfigModel):
    enabled: bool = True
    ep_size: int = 1
    moe_experts: list = Field([1], alias="num_experts")
    type: MoETypeEnum = MoETypeEnum.standard
    ep_mp_group: object = None
    ep_group: object = Field(None, alias="expert_group")


class QuantTypeEnum(str, Enum):
    asym = "asymmetric"
    sym = "symmetric"


class BaseQuantConfig(DeepSpeedConfigModel):
    enabled: bool = True
    num_bits: int = 8
    q_type: QuantTypeEnum = QuantTypeEnum.sym
    q_groups: int = 1


class WeightQuantConfig(BaseQuantConfig):
    enabled: bool = True
    quantized_initialization: Dict = {}
    post_init_quant: Dict = {}


class ActivationQuantConfig(BaseQuantConfig):
    enabled: bool = True


class QKVQuantConfig(DeepSpeedConfigModel):
    enabled: bool = True


class QuantizationConfig(DeepSpeedConfigModel):
    enabled: bool = True
    activation: ActivationQuantConfig = ActivationQuantConfig()
    weight: WeightQuantConfig = WeightQuantConfig()
    qkv: QKVQuantConfig = QKVQuantConfig()


class InferenceCheckpointConfig(DeepSpeedConfigModel):
    checkpoint_dir: Optional[str] = None
    save_mp_checkpoint_path: Optional[str] = None
    base_dir: Optional[str] = None


class DeepSpeedInferenceConfig(DeepSpeedConfigModel):
    replace_with_kernel_inject: bool = Field(False, alias="kernel_inject")
    dtype: object = torch.float16
    tensor_parallel: DeepSpeedTPConfig = Field({}, alias="tp")
    enable_cuda_graph: bool = False
    use_triton: bool = False        # accepted for config compatibility; there is no Triton path here
    triton_autotune: bool = False
    zero: DeepSpeedZeroConfig = {}
    triangular_masking: bool = Field(True, alias="tm")
    moe: Union[bool, DeepSpeedMoEConfig] = {}
    keep_module_on_host: bool = False
    quant: QuantizationConfig = {}
    checkpoint: Optional[Union[str, Dict]] = None
    base_dir: str = ""
    set_empty_params: bool = False
    save_mp_checkpoint_path: Optional[str] = None
    checkpoint_config: InferenceCheckpointConfig = Field({}, alias="ckpt_config")
    return_tuple: bool = True
    training_mp_size: int = 1
    replace_method: str = Field("auto", json_schema_extra={"deprecated": True})
    injection_policy: Optional[Dict] = Field(None, alias="injection_dict")
    injection_policy_tuple: Optional[tuple] = None
    config: Optional[Dict] = Field(None, alias="args")
    max_out_tokens: int = Field(1024, alias="max_tokens")
    min_out_tokens: int = Field(1, alias="min_tokens")
    transposed_mode: bool = False
    mp_size: int = Field(1, json_schema_extra={"deprecated": True, "new_param": "tensor_parallel.tp_size"})
    mpu: object = Field(None, json_schema_extra={"deprecated": True, "new_param": "tensor_parallel.mpu"})
    ep_size: int = Field(1, json_schema_extra={"deprecated": True, "new_param": "moe.ep_size"})
    ep_group: object = Field(None, alias="expert_group", json_schema_extra={"deprecated": True, "new_param": "moe.ep_group"})
    ep_mp_group: object = Field(None, alias="expert_mp_group", json_schema_extra={"deprecated": True, "new_param": "moe.ep_mp_group"})
    moe_experts: list = Field([1], json_schema_extra={"deprecated": True, "new_param": "moe.moe_experts"})
    moe_type: MoETypeEnum = Field(MoETypeEnum.standard, json_schema_extra={"deprecated": True, "new_param": "moe.type"})

    @field_validator("dtype", mode="before")
    @classmethod
    def _dtype(cls, v):
        return DtypeEnum.from_any(v).torch

    @field_validator("moe", mode="before")
    @classmethod
    def _moe(cls, v):
        if isinstance(v, bool):
            return DeepSpeedMoEConfig(enabled=v)
        return v
