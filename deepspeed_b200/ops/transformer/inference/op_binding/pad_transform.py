"""``PadTransformOp`` (reference ``ops/transformer/inference/op_binding/pad_transform.py``): ``[b, s, heads*d]`` -> ``[b, heads, s, d_padded]`` with the head dim padded to a multiple of 8 (diffusers attention)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class PadTransformOp(BaseOp):

    def forward(self, query, key, value, heads, add_padding=False):
        def tr(t):
            b, s, hd = t.shape
            t = t.view(b, s, heads, hd // heads).transpose(1, 2)
            pad = (-t.shape[-1]) % 8 if add_padding else 0
            return F.pad(t, (0, pad)).contiguous() if pad else t.contiguous()
        return tr(query), tr(key), tr(value)
