"""``res = x + y; hidden = RMSNorm(res)``.

Reference ``inference/v2/kernels/core_ops/cuda_rms_norm/rms_pre_norm.py``."""
import torch

from deepspeed_b200.ops.kernels import transformer_ops as T

from .rms_norm_base import CUDARMSNormBase


class CUDARMSPreNorm(CUDARMSNormBase):

    def __call__(self, z_res, z_hid, x_res, y_hid, gamma):
        hid, res = T.rms_norm(y_hid, gamma, self.epsilon, residual=x_res)
        z_res.copy_(res)
        z_hid.copy_(hid)
        return z_res, z_hid
