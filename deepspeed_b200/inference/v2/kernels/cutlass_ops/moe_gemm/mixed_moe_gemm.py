"""Grouped GEMM with weight-only-quantised expert weights.

Reference ``inference/v2/kernels/cutlass_ops/moe_gemm/mixed_moe_gemm.py``."""
import torch

from deepspeed_b200.inference.quantization.layers import maybe_quantized_linear
from deepspeed_b200.utils.types import ActivationFuncType

from ...ds_kernel import DSKernelBase, check_dtype


class MixedMoEGEMM(DSKernelBase):

    def __init__(self, fp_dtype, act_fn=ActivationFuncType.UNKNOWN, num_bits: int = 8) -> None:
        check_dtype(fp_dtype, "MixedMoEGEMM", allow_fp32=False)
        if num_bits not in (4, 8):
            raise ValueError("num_bits must be 4 or 8")
        self.act_fn, self.num_bits = act_fn, num_bits

    def __call__(self, ordered_output, ordered_input, weights, scales, cumsum_rows, biases=None) -> None:
        """``weights``: list of QuantizedWeight, one per expert."""
        ends = cumsum_rows.tolist() if torch.is_tensor(cumsum_rows) else list(cumsum_rows)
        s = 0
        for e, t in enumerate(ends):
            t = int(t)
            if t > s:
                ordered_output[s:t] = maybe_quantized_linear(ordered_input[s:t], weights[e], None if biases is None else biases[e])
            s = t
        return ordered_output
