"""``act(x + bias)`` in place (identity / relu / gelu / silu).

Reference ``inference/v2/kernels/core_ops/bias_activations/bias_activation.py``."""
import torch

from deepspeed_b200.ops.kernels import transformer_ops as T
from deepspeed_b200.utils.types import ActivationFuncType

from ...ds_kernel import DSKernelBase, check_dtype

_NAMES = {ActivationFuncType.UNKNOWN: None, ActivationFuncType.GELU: "gelu", ActivationFuncType.ReLU: "relu"}


class CUDABiasActivation(DSKernelBase):

    def __init__(self, channels: int, dtype, act_fn) -> None:
        check_dtype(dtype, "CUDABiasActivation")
        if channels % 8 != 0:
            raise ValueError("channels must be divisible by 8 (16-byte vector accesses)")
        if isinstance(act_fn, ActivationFuncType) and act_fn not in _NAMES:
            raise ValueError(f"Unsupported activation function {act_fn}; use the gated-activation kernel for gated types")
        self.act = _NAMES.get(act_fn, str(act_fn).lower() if act_fn is not None else None)
        self.act = None if self.act in ("identity", "none") else self.act

    def __call__(self, activation: torch.Tensor, bias: torch.Tensor = None) -> torch.Tensor:
        out = T.bias_act(activation, bias, act=self.act)
        if out.data_ptr() != activation.data_ptr():
            activation.copy_(out)
        return activation
