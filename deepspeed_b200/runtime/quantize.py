"""MoQ: progressive quantise-in-training (reference ``runtime/quantize.py:14 Quantizer``).  Every
``q_period`` steps the target precision of a group drops by one bit (period doubling each time, or driven by the
layer's Hessian eigenvalue), and weights are fake-quantised in place after the optimizer step with the
``dsb_fake_quantize`` kernel (symmetric / asymmetric, nearest / stochastic rounding), optionally blended with the
fp weights by ``quantize_real_ratio`` (mixed-fp16 quantisation)."""
import math

import torch

from deepspeed_b200.ops.quantizer import quantizer as Q
from deepspeed_b200.utils.logging import logger

TWO_D_PARAMS = 6


class Quantizer:

    def __init__(self, q_groups=1, q_mixed_fp16=False, q_change_ratio=0.01, q_type=0, q_rounding=0, q_verbose=False,
                 q_eigenvalue=False, use_quantizer_kernel=False, layer_num=0):
        self.q_groups = q_groups
        self.q_mixed_fp16 = q_mixed_fp16
        self.q_change_ratio = q_change_ratio
        self.q_type = q_type              # 0 symmetric, 1 asymmetric
        self.qsteps = 0
        self.quantize_real_ratio = 1.000
        self.q_verbose = q_verbose
        self.q_eigenvalue = q_eigenvalue
        self.use_quantizer_kernel = use_quantizer_kernel
        self.q_rounding = q_rounding      # 0 nearest, 1 stochastic
        self.layer_num = layer_num
        self.q_start_bits, self.q_target_bits, self.q_period = [], [], []

    def any_precision_switch(self):
        if self.layer_num == 0:
            return True
        for i in range(self.layer_num):
            if self.q_start_bits[i] != self.q_target_bits:
                if self.qsteps + (TWO_D_PARAMS * (self.layer_num if self.layer_num != 0 else 1)) >= self.q_period[i]:
                    return True
        return False

    def quantize(self, parameter_group, overflow, eigenvalue_enabled, block_eigenvalue=None):
        if overflow and not eigenvalue_enabled:
            return
        self.step()
        self.update_fp16_ratio()
        block_eigenvalue = block_eigenvalue or {}
        for i, group in enumerate(parameter_group):
            for p in group:
                if len(p.size()) > 1 and getattr(p, "start_bits", 0):
                    eig, layer_id = block_eigenvalue.get(id(p), (None, 0))
                    factor = 1 + math.floor(eig * 4) if eig is not None else None
                    p.data = self.compute_quantization(p, layer_id, factor)

    def step(self):
        self.qsteps += 1

    def update_fp16_ratio(self):
        if self.q_mixed_fp16:
            self.quantize_real_ratio = max(0.0, self.quantize_real_ratio - self.q_change_ratio)

    def compute_quantization(self, p, index=0, factor=None):
        if p.start_bits != p.target_bits and self.qsteps >= p.q_period:
            self.quantize_real_ratio = 1.0
            p.q_period = p.q_period * 2 * (factor if factor is not None else 1) if factor is not None else p.q_period * 2
            p.start_bits -= 1
            if self.q_verbose:
                logger.info(f"Quantization settings: current bit-precision = {p.start_bits}, step = {self.qsteps}, "
                            f"quantization period = {p.q_period}, index = {index}")
        assert p.start_bits >= p.target_bits, "Quantization bit is lower than target precision bits!"
        x = p.data
        groups = self.q_groups
        while x.numel() % groups:
            groups -= 1
        if p.start_bits >= 16:
            q = x
        elif p.start_bits == 1:
            flat = x.reshape(groups, -1).float()
            m = flat.abs().mean(dim=1, keepdim=True)
            q = (flat.sign() * m).reshape(x.shape).to(x.dtype)
        elif p.start_bits == 2:
            flat = x.reshape(groups, -1).float()
            thres = 0.7 * flat.abs().mean(dim=1, keepdim=True)
            mask = (flat.abs() > thres).float()
            alpha = (flat.abs() * mask).sum(1, keepdim=True) / mask.sum(1, keepdim=True).clamp(min=1)
            q = (alpha * flat.sign() * mask).reshape(x.shape).to(x.dtype)
        else:
            q = Q.fake_quantize(x.contiguous().view(-1), groups, p.start_bits,
                                Q.Symmetric if self.q_type == 0 else Q.Asymmetric, stochastic=self.q_rounding == 1,
                                seed=self.qsteps).view(x.shape)
        if self.q_mixed_fp16 and p.start_bits >= p.target_bits - 1 and self.quantize_real_ratio > 0:
            q = self.quantize_real_ratio * x + (1 - self.quantize_real_ratio) * q
        return q
