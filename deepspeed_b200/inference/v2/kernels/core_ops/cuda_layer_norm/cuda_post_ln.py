"""``LayerNorm(x + y)`` (post-LN residual).

Reference ``inference/v2/kernels/core_ops/cuda_layer_norm/cuda_post_ln.py``."""
import torch

from deepspeed_b200.ops.kernels import transformer_ops as T

from .cuda_fp_ln_base import CUDAFPLNBase


class CUDAFPPostLN(CUDAFPLNBase):

    def __call__(self, output_z, input_x, input_y, gamma, beta) -> torch.Tensor:
        out, _ = T.layer_norm(input_y, gamma, beta, self.epsilon, residual=input_x)
        output_z.copy_(out)
        return output_z
