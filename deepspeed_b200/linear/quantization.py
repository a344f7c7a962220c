"""``QuantizedParameter``: a frozen parameter stored FP8/FP6-quantised (``ops/fp_quantizer``), dequantised on use
(reference ``linear/quantization.py:18``)."""
import copy

import torch
from torch import nn
import torch.nn.functional as F

from deepspeed_b200.ops.fp_quantizer.quantize import FP_Quantize
from .config import QuantizationConfig


class QuantizedParameter(nn.Parameter):

    def __new__(cls, data=None, requires_grad=False, quantization_config: QuantizationConfig = None, quantizer=None, **state):
        """``state``: the remaining entries of another instance's ``__dict__`` -- HF clones parameters as
        ``type(p)(p.data, **p.__dict__)``; an already quantised payload is adopted as is."""
        if requires_grad:
            raise ValueError("requires_grad=True is not supported with QuantizedParameter")
        if data is None:
            data = torch.empty(0)
        self = torch.Tensor._make_subclass(cls, data, requires_grad)
        self.quantization_config = quantization_config or QuantizationConfig()
        self.quantizer = quantizer if quantizer is not None else FP_Quantize(group_size=self.quantization_config.group_size)
        self._orig_shape, self._orig_dtype, self._scale = None, None, None
        for k, v in state.items():
            setattr(self, k, v)
        self._ensure_quantized(self)
        return self

    def _ensure_quantized(self, tensor: torch.Tensor):
        # quantise once the payload sits on the accelerator (or immediately on host-only runs) and is still fp
        if self._scale is not None or tensor.numel() == 0 or not tensor.is_floating_point() or tensor.dtype == torch.float64:
            return
        if tensor.device.type == "cuda" or not torch.cuda.is_available():
            self._orig_shape, self._orig_dtype = tensor.shape, tensor.dtype
            with torch.no_grad():
                q, scale = self.quantizer.quantize(tensor.data.contiguous(), q_bits=self.quantization_config.q_bits,
                                                   q_mantisa_bits=self.quantization_config.mantissa_bits,
                                                   return_meta_tensor=True)
            self._scale = scale
            tensor.data = q
            self.quantizer.orig_shape, self.quantizer.orig_dtype = self._orig_shape, self._orig_dtype

    def dequantized(self) -> torch.Tensor:
        if self._scale is None:
            return self.data
        self.quantizer.orig_shape, self.quantizer.orig_dtype = self._orig_shape, self._orig_dtype
        with torch.no_grad():
            return self.quantizer.dequantize(self.data, q_bits=self.quantization_config.q_bits,
                                             q_mantisa_bits=self.quantization_config.mantissa_bits,
                                             scale=self._scale).view(self._orig_shape).to(self._orig_dtype)

    def offload(self, revert=False):
        dev = ("cuda" if torch.cuda.is_available() else "cpu") if revert else "cpu"
        self.data = self.data.to(dev)
        if self._scale is not None:
            self._scale = self._scale.to(dev)

    def __getstate__(self):
        st = dict(self.__dict__)
        st["data"], st["requires_grad"] = self.data, self.requires_grad
        return st

    def __setstate__(self, st):
        self.__dict__.update({k: v for k, v in st.items() if k not in ("data", "requires_grad")})
        self.data, self.requires_grad = st["data"], st["requires_grad"]

    def __deepcopy__(self, memo):
        new = type(self).__new__(type(self), quantization_config=self.quantization_config, quantizer=self.quantizer)
        st = self.__getstate__()
        st["data"] = copy.deepcopy(st["data"])
        st["_scale"] = copy.deepcopy(st.get("_scale"))
        new.__setstate__(st)
        return new

    def __copy__(self):
        new = type(self).__new__(type(self), quantization_config=self.quantization_config, quantizer=self.quantizer)
        new.__setstate__(self.__getstate__())
        return new

    def cuda(self, device=None, non_blocking=False):
        return self.to(device="cuda" if device is None else device, non_blocking=non_blocking)

    def to(self, *args, **kwargs):
        t = super().to(*args, **kwargs)
        out = QuantizedParameter.__new__(QuantizedParameter, t.data if self._scale is not None else t,
                                         quantization_config=self.quantization_config, quantizer=self.quantizer) \
            if self._scale is None else self.__copy__()
        if self._scale is not None:
            out.data = t.data
            out._scale = self._scale.to(t.device)
        return out


_Linear = nn.Linear


class QuantizedLinear(_Linear):
    """nn.Linear whose weight is a ``QuantizedParameter`` (dequantised per forward)."""

    def __init__(self, input_dim: int, output_dim: int, bias: bool = False, quantization_config: QuantizationConfig = None,
                 dtype=torch.bfloat16):
        super().__init__(input_dim, output_dim, bias=bias, dtype=dtype)
        assert dtype == torch.bfloat16, "currently only supports bfloat16 dtype"
        self.weight = QuantizedParameter(self.weight.data, quantization_config=quantization_config)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return F.linear(input, self.weight.dequantized(), self.bias)
