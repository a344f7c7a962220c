"""Fused MoE inference layer (reference ``ops/transformer/inference/moe_inference.py:163 DeepSpeedMoEInference``):
attention block + top-k routed experts (expert parallel over ``ep_group``, optional residual-MoE mixing)."""
import torch
from torch import nn
import torch.nn.functional as F

from deepspeed_b200 import comm as dist
from deepspeed_b200.ops.kernels import moe_ops as M
from .config import DeepSpeedInferenceConfig
from .ds_transformer import DeepSpeedTransformerInference, _act


class DeepSpeedMoEInferenceConfig(DeepSpeedInferenceConfig):

    def __init__(self, *args, moe_experts=1, k=1, capacity_factor=1.0, eval_capacity_factor=1.0, min_capacity=1,
                 noisy_gate_policy=None, drop_tokens=True, use_rts=False, global_experts=1, mlp_type="standard",
                 scale_attn_by_inverse_layer_idx=False, **kw):
        super().__init__(*args, **kw)
        self.moe_experts = moe_experts
        self.k = k
        self.capacity_factor = capacity_factor
        self.eval_capacity_factor = eval_capacity_factor
        self.min_capacity = min_capacity
        self.noisy_gate_policy = noisy_gate_policy
        self.drop_tokens = drop_tokens
        self.use_rts = use_rts
        self.global_experts = global_experts
        self.mlp_type = mlp_type
        self.scale_attn_by_inverse_layer_idx = scale_attn_by_inverse_layer_idx


class DeepSpeedMLPFunction(torch.autograd.Function):
    """Inference-only expert MLP: ``act(x W1ᵀ + b1) W2ᵀ`` (+ all-reduce over the expert model-parallel group); the output bias
    is added by the caller after the reduction (reference ``moe_inference.py:104``)."""

    @staticmethod
    def forward(ctx, input, inter_w, inter_b, config, output_b, output_w, q_scales=None, q_groups=1, merge_count=1,
                mp_group=None, async_op=False):
        h = _act(F.linear(input, inter_w, inter_b), config.mlp_act_func_type)
        out = F.linear(h, output_w)
        if mp_group is not None and dist.is_initialized() and dist.get_world_size(group=mp_group) > 1:
            dist.all_reduce(out, group=mp_group, async_op=async_op)
        return out + output_b

    @staticmethod
    def backward(ctx, grad_output):
        raise RuntimeError("You are running with DeepSpeed Inference mode. Please switch to Training mode for running "
                           "backward!")


class DeepSpeedMoEMLP(nn.Module):
    """One expert's FFN, its hidden dimension sharded over the expert model-parallel group (reference ``:138``)."""

    def __init__(self, config, q_scales=None, q_groups=1, merge_count=1, mlp_extra_grouping=False, mp_group=None):
        super().__init__()
        self.config = config
        dt = torch.float16 if getattr(config, "fp16", False) else (torch.bfloat16 if getattr(config, "bf16", False) else torch.float32)
        dt = getattr(config, "dtype", dt) or dt
        mp = dist.get_world_size(group=mp_group) if (mp_group is not None and dist.is_initialized()) else 1
        inter = config.intermediate_size // mp
        p = lambda *shape: nn.Parameter(torch.empty(*shape, dtype=dt), requires_grad=False)
        self.attn_nw, self.attn_nb = p(config.hidden_size), p(config.hidden_size)
        self.inter_w, self.inter_b = p(inter, config.hidden_size), p(inter)
        self.output_w, self.output_b = p(config.hidden_size, inter), p(config.hidden_size)
        self.q_scales, self.q_groups, self.merge_count, self.mp_group = q_scales, q_groups, merge_count, mp_group

    def forward(self, input, async_op=False):
        return DeepSpeedMLPFunction.apply(input, self.inter_w, self.inter_b, self.config, self.output_b, self.output_w,
                                          self.q_scales, self.q_groups, self.merge_count, self.mp_group, async_op)


class DeepSpeedMoEInference(DeepSpeedTransformerInference):

    def __init__(self, config, mp_group=None, ep_group=None, expert_mp_group=None, quantize_scales=None,
                 quantize_groups=1, merge_count=1, mlp_extra_grouping=False):
        super().__init__(config, mp_group)
        c = config
        self.ep_group = ep_group
        self.ep_size = dist.get_world_size(ep_group) if ep_group is not None else 1
        n_local = c.moe_experts if isinstance(c.moe_experts, int) else c.moe_experts[0]
        self.n_local = n_local
        self.n_global = max(c.global_experts, n_local * self.ep_size)
        dt = self.inter_w.dtype
        h, i = c.hidden_size, c.intermediate_size
        p = lambda *s: nn.Parameter(torch.empty(*s, dtype=dt), requires_grad=False)
        self.gate_w = p(self.n_global, h)
        self.expert_inter_w, self.expert_inter_b = p(n_local, i, h), p(n_local, i)
        self.expert_out_w, self.expert_out_b = p(n_local, h, i), p(n_local, h)
        if c.mlp_type == "residual":
            self.coef_w, self.coef_b = p(2, h), p(2)

    def _experts(self, x2):
        c = self.config
        Tn, H = x2.shape
        ids, w, _ = M.top_k_gating(F.linear(x2, self.gate_w).float(), c.k, normalize=c.k > 1)
        out = torch.zeros(Tn, H, dtype=torch.float32, device=x2.device)
        ep_rank = dist.get_rank(self.ep_group) if self.ep_group is not None else 0
        # only the tokens routed to an expert visit it: sort the (token, slot) pairs by expert once, then run each local
        # expert on its contiguous run
        flat_e = ids.reshape(-1)
        order = torch.argsort(flat_e, stable=True)
        tok = (order // ids.shape[1])
        counts = torch.bincount(flat_e, minlength=self.n_global)
        starts = torch.cumsum(counts, 0) - counts
        wflat = w.reshape(-1)[order]
        lo, hi = ep_rank * self.n_local, (ep_rank + 1) * self.n_local
        bounds = torch.stack([starts[lo:hi], counts[lo:hi]], 1).tolist()
        for le, (s0, n) in enumerate(bounds):
            if n == 0:
                continue
            rows = tok[s0:s0 + n]
            ye = F.linear(_act(F.linear(x2[rows], self.expert_inter_w[le], self.expert_inter_b[le]), c.mlp_act_func_type),
                          self.expert_out_w[le], self.expert_out_b[le])
            out.index_add_(0, rows, ye.float() * wflat[s0:s0 + n, None])
        if self.ep_size > 1:
            dist.all_reduce(out, group=self.ep_group)
        return out.to(x2.dtype)

    def forward(self, input, input_mask=None, attention_mask=None, head_mask=None, layer_past=None, get_key_value=False,
                get_present=False, encoder_output=None, enc_dec_attn_mask=None, encoder_hidden_states=None,
                encoder_attention_mask=None, use_cache=False, output_attentions=False, **kw):
        c = self.config
        x = input
        if not use_cache and layer_past is None:
            self.seen = 0
        a_in = self._norm(x, self.norm_w, self.norm_b) if c.pre_layer_norm else x
        a = self._reduce(F.linear(self._attn(a_in, attention_mask if attention_mask is not None else input_mask),
                                  self.attn_ow))
        f_in, residual = self._norm(a + self.attn_ob, self.attn_nw, self.attn_nb, residual=x)
        B, S, H = f_in.shape
        moe = self._experts(f_in.reshape(-1, H)).view(B, S, H)
        if c.mlp_type == "residual":
            dense = F.linear(_act(F.linear(f_in, self.inter_w, self.inter_b), c.mlp_act_func_type), self.output_w,
                             self.output_b)
            coef = torch.softmax(F.linear(f_in, self.coef_w, self.coef_b), dim=-1)
            moe = dense * coef[..., 0:1] + moe * coef[..., 1:]
        out = residual + moe
        return (out, None) if c.return_tuple else out
