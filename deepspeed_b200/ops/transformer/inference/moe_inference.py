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
        for le in range(self.n_local):
            ge = ep_rank * self.n_local + le
            we = (w * (ids == ge)).sum(-1, keepdim=True)
            ye = F.linear(_act(F.linear(x2, self.expert_inter_w[le], self.expert_inter_b[le]), c.mlp_act_func_type),
                          self.expert_out_w[le], self.expert_out_b[le])
            out.addcmul_(ye.float(), we)
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
