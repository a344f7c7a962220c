"""Module replacement entry points (reference ``module_inject/replace_module.py``:
``replace_transformer_layer :183``, ``generic_injection :88``, ``revert_transformer_layer``).

Kernel injection for causal LMs is done at whole-model granularity by ``inference.engine.InferenceEngine``
(weights re-packed into the ragged fused model).  ``replace_transformer_layer`` keeps the reference entry
point for callers that drive it directly: it applies AutoTP sharding and optional weight quantisation.
``generic_injection`` covers the diffusers-style path: attention blocks get the fused SDPA attention.
"""
import torch
from torch import nn

from .auto_tp import AutoTP


def replace_transformer_layer(orig_layer_impl, model, checkpoint_dict=None, config=None, model_config=None):
    tp = config.tensor_parallel.tp_size if config is not None else 1
    if tp > 1:
        from deepspeed_b200.utils import groups
        if groups.ranks_of("tp") is None:
            groups._init_tp_mesh_device(tensor_model_parallel_size=tp)
        AutoTP(model, mp_group=groups.get_tensor_model_parallel_group(), mp_size=tp).replace()
    if config is not None and config.quant.enabled and config.quant.weight.post_init_quant:
        from deepspeed_b200.inference.quantization import _init_group_wise_weight_quantization
        _init_group_wise_weight_quantization(model, {"weight_quantization": {"post_init_quant":
                                                                             config.quant.weight.post_init_quant}})
    return model


class _FusedSelfAttention(nn.Module):
    """Replacement for diffusers ``CrossAttention`` / ``Attention`` blocks: packed QKV GEMM + flash SDPA
    (reference ``DeepSpeedDiffusersAttention``)."""

    def __init__(self, attn):
        super().__init__()
        self.heads = attn.heads
        self.to_q, self.to_k, self.to_v = attn.to_q, attn.to_k, attn.to_v
        self.to_out = attn.to_out[0] if isinstance(attn.to_out, (nn.ModuleList, nn.Sequential)) else attn.to_out
        self._packed = None

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        B, S, _ = hidden_states.shape
        if encoder_hidden_states is None:
            if self._packed is None:
                self._packed = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], 0).detach()
            q, k, v = torch.nn.functional.linear(hidden_states, self._packed).chunk(3, dim=-1)
        else:
            q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        h = self.heads
        q, k, v = (t.reshape(B, -1, h, t.shape[-1] // h).transpose(1, 2) for t in (q, k, v))
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask)
        return self.to_out(o.transpose(1, 2).reshape(B, S, -1))


def generic_injection(module, dtype=None, enable_cuda_graph=True):
    """Swap diffusers attention blocks for the fused attention; returns the number replaced."""
    n = 0
    for name, sub in list(module.named_modules()):
        for cname, child in list(sub.named_children()):
            if type(child).__name__ in ("CrossAttention", "Attention") and all(hasattr(child, a) for a in
                                                                             ("to_q", "to_k", "to_v", "to_out", "heads")):
                new = _FusedSelfAttention(child)
                if dtype is not None:
                    new = new.to(dtype)
                setattr(sub, cname, new)
                n += 1
    return n


def revert_transformer_layer(orig_layer_impl, model, config, preln=False):
    raise NotImplementedError("revert is not supported: keep a reference to the original module instead")
