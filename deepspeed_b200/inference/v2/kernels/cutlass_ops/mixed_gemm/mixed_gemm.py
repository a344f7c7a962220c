"""Weight-only-quantised GEMM: int8 / int4 weights + per-group scales, fp16/bf16 activations (fused dequant GEMV at decode sizes, dequantise + tensor-core GEMM otherwise).

Reference ``inference/v2/kernels/cutlass_ops/mixed_gemm/mixed_gemm.py``."""
import torch

from deepspeed_b200.inference.quantization.layers import maybe_quantized_linear, quantize_weight
from deepspeed_b200.utils.types import ActivationFuncType

from ...ds_kernel import DSKernelBase, check_dtype


class MixedGEMM(DSKernelBase):
    supported_dtypes = [torch.float16, torch.bfloat16]

    def __init__(self, fp_dtype, act_fn=ActivationFuncType.UNKNOWN, num_bits: int = 8) -> None:
        check_dtype(fp_dtype, "MixedGEMM", allow_fp32=False)
        if num_bits not in (4, 8):
            raise ValueError("num_bits must be 4 or 8")
        self.num_bits = num_bits
        self.act_fn = act_fn

    def quantize(self, weight: torch.Tensor, group_size: int = 128):
        return quantize_weight(weight, "int8" if self.num_bits == 8 else "int4", group_size)

    def __call__(self, output, hidden_states, weights, scales=None, biases=None) -> None:
        """``weights``: a QuantizedWeight (codes + scales travel together; ``scales`` is accepted for signature parity)."""
        res = maybe_quantized_linear(hidden_states, weights, biases)
        if self.act_fn in (ActivationFuncType.GELU, ActivationFuncType.ReLU):
            res = torch.nn.functional.gelu(res) if self.act_fn == ActivationFuncType.GELU else torch.relu(res)
        output.copy_(res.reshape(output.shape))
        return output
