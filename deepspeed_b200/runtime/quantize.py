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
        if p.start_bits >= 16:
            return x
        if p.start_bits == 1:
            q = self.quantize_binary(x)
        elif p.start_bits == 2:
            q = self.quantize_tenary(x)
        else:
            q = self.quantize_highbit(x, p.start_bits)
        return self.mixed_fp16_quantize(x, q, index, start_bits=p.start_bits, target_bits=p.target_bits)

    # ---- the three quantisers (reference ``quantize.py:78-127``) -----------------------------------------------------
    def _groups_for(self, x):
        groups = self.q_groups
        while x.numel() % groups:
            groups -= 1
        return groups

    def quantize_highbit(self, inputs, num_bits):
        """Uniform fake quantisation to ``num_bits`` (≥3) per group on the native quantiser kernel."""
        return Q.fake_quantize(inputs.contiguous().view(-1), self._groups_for(inputs), num_bits,
                               Q.Symmetric if self.q_type == 0 else Q.Asymmetric, stochastic=self.q_rounding == 1,
                               seed=self.qsteps).view(inputs.shape)

    def quantize_tenary(self, inputs):
        """Ternary {-α, 0, +α}: threshold 0.7·mean|w| per group, α the mean magnitude of the surviving weights."""
        flat = inputs.reshape(self._groups_for(inputs), -1).float()
        thres = 0.7 * flat.abs().mean(dim=1, keepdim=True)
        mask = (flat.abs() > thres).float()
        alpha = (flat.abs() * mask).sum(1, keepdim=True) / mask.sum(1, keepdim=True).clamp(min=1)
        return (alpha * flat.sign() * mask).reshape(inputs.shape).to(inputs.dtype)

    def quantize_binary(self, inputs):
        """Binary sign(w)·mean|w| per group."""
        flat = inputs.reshape(self._groups_for(inputs), -1).float()
        return (flat.sign() * flat.abs().mean(dim=1, keepdim=True)).reshape(inputs.shape).to(inputs.dtype)

    def mixed_fp16_quantize(self, input, input_q, index, start_bits=None, target_bits=None):
        """Blend the real and the quantised weights while the schedule is within one bit of the target
        (``quantize_real_ratio`` decays to 0 through ``update_fp16_ratio``)."""
        if start_bits is None:
            start_bits = self.q_start_bits[index] if hasattr(self, "q_start_bits") else None
            target_bits = getattr(self, "q_target_bits", None)
        near = start_bits is not None and target_bits is not None and start_bits >= target_bits - 1
        if self.q_mixed_fp16 and near and self.quantize_real_ratio > 0:
            return self.quantize_real_ratio * input + (1 - self.quantize_real_ratio) * input_q
        return input_q
