"""Stand-alone fused MLP block of the v1 inference layer (reference ``ops/transformer/inference/ds_mlp.py``):
(residual + bias) -> norm -> GEMM -> activation -> GEMM -> residual combine."""
import torch
from torch import nn

from deepspeed_b200 import comm as dist

from .op_binding import MLPGemmOp, ResidualAddOp


class DeepSpeedMLP(nn.Module):

    def __init__(self, config, mp_group=None, q_scales=None, q_groups=1, merge_count=1, mlp_extra_grouping=False):
        super().__init__()
        self.config = config
        c = config
        tp = c.mp_size
        dt = c.dtype if c.dtype in (torch.float16, torch.bfloat16, torch.float32) else torch.float16
        from .ds_transformer import _is_gated
        rows = (2 if _is_gated(c.mlp_act_func_type) else 1) * (c.intermediate_size // tp)
        p = lambda *s: nn.Parameter(torch.empty(*s, dtype=dt), requires_grad=False)
        self.attn_nw, self.attn_nb = p(c.hidden_size), p(c.hidden_size)
        self.inter_w, self.inter_b = p(rows, c.hidden_size), p(rows)
        self.output_w, self.output_b = p(c.hidden_size, c.intermediate_size // tp), p(c.hidden_size)
        self.mp_group = mp_group
        self.mlp_gemm_func, self.residual_add_func = MLPGemmOp(c), ResidualAddOp(c)

    def forward(self, input, residual, residual_norm=None, bias=None):
        """``input``: attention output (pre-bias), ``bias``: attention output bias, ``residual``: the layer input."""
        out, res = self.mlp_gemm_func(input, residual, self.inter_w, self.output_w, bias, self.inter_b, self.attn_nw, self.attn_nb)
        if self.mp_group is not None and dist.get_world_size(self.mp_group) > 1:
            dist.inference_all_reduce(out, group=self.mp_group)
        return self.residual_add_func(out, res, add_bias=True, final_bias=self.output_b)
