"""Plain RMSNorm.

Reference ``inference/v2/kernels/core_ops/cuda_rms_norm/rms_norm.py``."""
import torch

from deepspeed_b200.ops.kernels import transformer_ops as T

from .rms_norm_base import CUDARMSNormBase


class CUDARMSNorm(CUDARMSNormBase):

    def __call__(self, output_z: torch.Tensor, input_x: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
        output_z.copy_(T.rms_norm(input_x, gamma, self.epsilon))
        return output_z
