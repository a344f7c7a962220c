"""``act(gate + b_g) * (up + b_u)`` over a ``[gate | up]`` stacked input (SwiGLU / GEGLU / ReGLU).

Reference ``inference/v2/kernels/core_ops/gated_activations/gated_activation.py``."""
import torch

from deepspeed_b200.ops.kernels import transformer_ops as T
from deepspeed_b200.utils.types import ActivationFuncType

from ...ds_kernel import DSKernelBase, check_dtype


class CUDAGatedActivation(DSKernelBase):
    supported_act_fns = [ActivationFuncType.GATED_GELU, ActivationFuncType.GATED_SILU]

    def __init__(self, channels: int, fp_dtype, act_fn) -> None:
        check_dtype(fp_dtype, "CUDAGatedActivation")
        if act_fn not in self.supported_act_fns and str(act_fn).lower() not in ("silu", "gelu", "relu"):
            raise ValueError(f"Unsupported activation function {act_fn}")
        if channels % 8 != 0:
            raise ValueError("channels must be divisible by 8")
        self.act = "silu" if act_fn == ActivationFuncType.GATED_SILU or "silu" in str(act_fn).lower() else (
            "relu" if "relu" in str(act_fn).lower() else "gelu")

    def __call__(self, output: torch.Tensor, input: torch.Tensor, bias: torch.Tensor = None) -> None:
        x = input if bias is None else input + bias
        output.copy_(T.gated_act(x.contiguous(), act=self.act))
        return output
