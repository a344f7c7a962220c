"""``GatedActivationOp`` (reference ``ops/transformer/inference/op_binding/gated_activation.py``): ``act(x[..., :h] + b) * (x[..., h:] + b)`` for GEGLU / SwiGLU (diffusers, llama)."""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels import misc_ops as M  # noqa: F401
from deepspeed_b200.ops.kernels import transformer_ops as T  # noqa: F401

from .base import BaseOp


class GatedActivationOp(BaseOp):

    def forward(self, activation: torch.Tensor, bias: torch.Tensor, activation_func_type):
        name = "silu" if "silu" in str(activation_func_type).lower() or str(int(activation_func_type)) == "4" else "gelu"
        x = activation if bias is None else activation + bias
        # the kernel expects [gate | up]; diffusers GEGLU stores [value | gate]
        up, gate = x.chunk(2, dim=-1)
        return T.gated_act(torch.cat([gate, up], dim=-1).contiguous(), act=name)
