"""Weight-only FP6 linear (FP6-LLM role): weights stored as packed fp6 + per-group scales, activations fp16/bf16.

Reference ``inference/v2/kernels/core_ops/cuda_linear/cuda_linear.py``."""
import torch

from deepspeed_b200.inference.quantization.layers import maybe_quantized_linear, quantize_weight

from ...ds_kernel import DSKernelBase


class CUDAWf6Af16Linear(DSKernelBase):

    def __init__(self):
        pass

    @staticmethod
    def quantize(weight: torch.Tensor, group_size: int = 128):
        """-> QuantizedWeight holding the fp6 codes + scales."""
        return quantize_weight(weight, "fp6", group_size)

    def __call__(self, output, hidden_states, weights_2bit=None, weights_4bit=None, scale=None, out_channels=None, tokens=None,
                 in_channels=None, qweight=None) -> torch.Tensor:
        """``qweight``: a QuantizedWeight from :meth:`quantize` (the reference's split 2-bit / 4-bit planes are one packed
        tensor here)."""
        qw = qweight if qweight is not None else weights_2bit
        res = maybe_quantized_linear(hidden_states, qw)
        output.copy_(res.reshape(output.shape))
        return output
