"""``InferenceParameter``: a tensor that carries auxiliary tensors (quantisation scales, ...) with it through ``.to()``
and flattening (reference ``inference/v2/inference_parameter.py``)."""
from typing import Dict

import torch

CORE_PARAM = "_ds_core_param_key"
STR_TO_DTYPE = {"torch.float32": torch.float32, "torch.float64": torch.float64, "torch.float16": torch.float16,
                "torch.bfloat16": torch.bfloat16, "torch.int64": torch.int64, "torch.int32": torch.int32,
                "torch.int16": torch.int16, "torch.int8": torch.int8, "torch.uint8": torch.uint8, "torch.bool": torch.bool}


class InferenceParameter(torch.Tensor):

    @staticmethod
    def __new__(cls, tensor, *args, **kwargs):
        new = super().__new__(cls, tensor, *args, **kwargs)
        if hasattr(tensor, "_aux_attrs"):
            new._aux_attrs = tensor._aux_attrs
        return new

    def to(self, *args, **kwargs):
        new = InferenceParameter(super().to(*args, **kwargs))
        aux = getattr(self, "_aux_attrs", None)
        if aux is not None:
            new._aux_attrs = {}
            for name, t in aux.items():
                moved = t.to(*args, **kwargs) if not t.is_floating_point() or "dtype" not in kwargs else t.to(
                    *[a for a in args if not isinstance(a, torch.dtype)], **{k: v for k, v in kwargs.items() if k != "dtype"})
                new._aux_attrs[name] = moved
                setattr(new, name, moved)
        return new

    @classmethod
    def initialize(cls, core_param: torch.Tensor, **kwargs) -> "InferenceParameter":
        """``core_param`` + named auxiliary tensors (reachable as attributes)."""
        param = InferenceParameter(core_param)
        param._aux_attrs = dict(kwargs)
        for name, t in kwargs.items():
            if hasattr(param, name):
                raise ValueError(f"Attribute {name} already exists on param.")
            if not isinstance(t, torch.Tensor):
                raise ValueError(f"Attribute {name} must be a tensor.")
            setattr(param, name, t)
        return param

    @classmethod
    def initialize_raw(cls, **kwargs) -> "InferenceParameter":
        """From a dict that holds the core tensor under ``CORE_PARAM`` (the flattened-model restore path)."""
        if CORE_PARAM not in kwargs:
            raise ValueError(f"Must provide core parameter, with key {CORE_PARAM}.")
        core = kwargs.pop(CORE_PARAM)
        return cls.initialize(core, **kwargs)

    @property
    def aux_attrs(self) -> Dict[str, torch.Tensor]:
        return getattr(self, "_aux_attrs", {})
