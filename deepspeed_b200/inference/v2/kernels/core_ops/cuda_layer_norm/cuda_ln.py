"""Plain LayerNorm.

Reference ``inference/v2/kernels/core_ops/cuda_layer_norm/cuda_ln.py``."""
import torch

from deepspeed_b200.ops.kernels import transformer_ops as T

from .cuda_fp_ln_base import CUDAFPLNBase


class CUDAFPLN(CUDAFPLNBase):

    def __call__(self, output_z: torch.Tensor, input_x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
        output_z.copy_(T.layer_norm(input_x, gamma, beta, self.epsilon))
        return output_z
