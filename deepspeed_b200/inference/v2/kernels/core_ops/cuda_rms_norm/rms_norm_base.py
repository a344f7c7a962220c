"""Shared validation of the RMSNorm kernels.

Reference ``inference/v2/kernels/core_ops/cuda_rms_norm/rms_norm_base.py``."""
import torch

from ...ds_kernel import DSKernelBase, check_dtype


class CUDARMSNormBase(DSKernelBase):
    supported_dtypes = [torch.float16, torch.bfloat16, torch.float32]

    def __init__(self, channels: int, fp_dtype, epsilon: float = 1e-5):
        check_dtype(fp_dtype, type(self).__name__)
        if channels * torch.empty(0, dtype=fp_dtype).element_size() % 16 != 0:
            raise ValueError("channels must be divisible by 16 bytes")
        self.epsilon = epsilon

    def __call__(self, *a, **k):
        raise NotImplementedError
