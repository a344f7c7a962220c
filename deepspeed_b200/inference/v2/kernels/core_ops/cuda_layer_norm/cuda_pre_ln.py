"""``res = x + y; hidden = LayerNorm(res)`` (pre-LN residual, both returned).

Reference ``inference/v2/kernels/core_ops/cuda_layer_norm/cuda_pre_ln.py``."""
import torch

from deepspeed_b200.ops.kernels import transformer_ops as T

from .cuda_fp_ln_base import CUDAFPLNBase


class CUDAFPPreLN(CUDAFPLNBase):

    def __call__(self, z_res, z_hid, x_res, y_hid, gamma, beta):
        hid, res = T.layer_norm(y_hid, gamma, beta, self.epsilon, residual=x_res)
        z_res.copy_(res)
        z_hid.copy_(hid)
        return z_res, z_hid
