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


def _inject_layers(model, config, mp_group=None, mp_size=1, device=None):
    """Swap every transformer layer a policy recognises for its fused counterpart; returns the number replaced."""
    from .replace_policy import policy_for, policy_to_ds_container
    from .policy import TransformerPolicy
    TransformerPolicy.hf_model_config = getattr(model, "config", None)
    count = 0

    def walk(parent):
        nonlocal count
        for name, child in list(parent.named_children()):
            pol_cls = policy_for(child)
            if pol_cls is None:
                walk(child)
                continue
            policy = pol_cls(child, inference=True)
            cont = policy_to_ds_container[pol_cls](policy=policy, config=config, model_config=TransformerPolicy.hf_model_config,
                                                   layer_id=count, child=child)
            cont.set_tensor_parallel_config(mp_size, mp_group)
            setattr(parent, name, cont.build(device=device))
            count += 1

    walk(model)
    return count


def replace_transformer_layer(orig_layer_impl, model, checkpoint_dict=None, config=None, model_config=None):
    """``config.replace_with_kernel_inject``: per-layer fused kernels through the policy/container registry; otherwise
    AutoTP sharding.  Optional post-init weight quantisation in both cases."""
    tp = config.tensor_parallel.tp_size if config is not None else 1
    group = None
    if tp > 1:
        from deepspeed_b200.utils import groups
        if groups.ranks_of("tp") is None:
            groups._init_tp_mesh_device(tensor_model_parallel_size=tp)
        group = groups.get_tensor_model_parallel_group()
    if config is not None and getattr(config, "replace_with_kernel_inject", False):
        n = _inject_layers(model, config, mp_group=group, mp_size=tp)
        if checkpoint_dict is not None:
            from .load_checkpoint import load_model_with_checkpoint
            load_model_with_checkpoint(model, checkpoint_dict, mp_group=group, mp_size=tp)
        if n == 0:
            raise ValueError("replace_with_kernel_inject: no injection policy matches this model; use AutoTP "
                             "(replace_with_kernel_inject=False) or provide an injection_policy")
    elif tp > 1:
        AutoTP(model, mp_group=group, mp_size=tp).replace()
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
    """Inverse of kernel injection: rebuild ``orig_layer_impl(config)`` layers and copy the fused weights back in the
    family's original layout.  Supported for the families whose policy can be written to (HF BERT-style encoders);
    other families should keep a reference to the original module."""
    from .containers.base import InjectedLayer
    from .replace_policy import policy_for
    n = 0
    for parent in list(model.modules()):
        for name, child in list(parent.named_children()):
            if not isinstance(child, InjectedLayer):
                continue
            new = orig_layer_impl(config)
            pol_cls = policy_for(new)
            if pol_cls is None:
                raise NotImplementedError(f"no policy can write into {type(new).__name__}")
            pol = pol_cls(new)
            f = child.fused
            qkvw, qkvb, ow, ob = pol.attention()
            w1, b1, w2, b2 = pol.mlp()
            anw, anb, inw, inb = pol.layernorm()
            import torch
            with torch.no_grad():
                # q/k/v are separate parameters in the original layer: the policy's packed view is a copy, so write
                # through the original modules
                a = new.attention
                for lin, (wt, bt) in zip((a.self.query, a.self.key, a.self.value),
                                         zip(f.attn_qkvw.chunk(3, 0), f.attn_qkvb.chunk(3, 0))):
                    lin.weight.copy_(wt)
                    lin.bias.copy_(bt)
                for dst, src in ((ow, f.attn_ow), (ob, f.attn_ob), (w1, f.inter_w), (b1, f.inter_b), (w2, f.output_w),
                                 (b2, f.output_b), (anw, f.attn_nw), (anb, f.attn_nb), (inw, f.norm_w), (inb, f.norm_b)):
                    if dst is not None:
                        dst.copy_(src)
            setattr(parent, name, new.to(f.attn_ow.device, f.attn_ow.dtype))
            n += 1
    return model


class GroupQuantizer:
    """Symmetric int8 group quantiser applied to weights as they are sharded into the fused layers (reference
    ``module_inject/replace_module.py:GroupQuantizer``)."""

    def __init__(self, q_int8=True, group_size=1, num_bits=8, num_groups=0):
        self.q_int8, self.group_size, self.num_bits, self.num_groups = q_int8, group_size, num_bits, num_groups

    def quantize(self, inputs, qkv=True, count=1, parallel_dim=0):
        if not self.q_int8 or not qkv:
            inputs = nn.Parameter(inputs, requires_grad=False)
            inputs.scale = torch.empty(1)
            return inputs
        q_range = 2**self.num_bits
        groups = self.num_groups if self.num_groups > 0 else max(1, inputs.shape[0] // self.group_size)
        flat = inputs.detach().float().reshape(groups, -1)
        bound = torch.maximum(flat.amax(1), flat.amin(1).abs())
        scale = q_range / (2 * bound + 1e-5)
        q = (flat * scale[:, None]).round().clamp(-q_range // 2, q_range // 2 - 1).reshape(inputs.shape).to(torch.int8)
        out = nn.Parameter(q, requires_grad=False)
        # the fused kernels consume INVERSE scales, one row per tensor-parallel slice of the weight
        parts = [1.0 / s for s in scale.reshape(count, -1)] if count > 1 else [1.0 / scale]
        out.scale = torch.cat([p.reshape(1, -1) for p in parts], dim=0).reshape(-1).unsqueeze(0)
        return out


def get_transformer_name(replaced_module):
    """Dotted path of the ``ModuleList`` holding the transformer layers (``"transformer.h"``, ``"model.layers"`` ...): the
    first child that contains a ModuleList of (injected or original) layers."""
    for n, c in replaced_module.named_children():
        for name, child in c.named_children():
            if isinstance(child, nn.ModuleList) and len(child) > 0:
                return f"{n}.{name}"
    return ""


def skip_level_0_prefix(model, state_dict):
    """Checkpoints of BLOOM / OPT drop the top-level module name from their keys -- unless the keys start with ``model.``."""
    import re
    if state_dict is not None and any(re.match(r"^model[.]", k) for k in state_dict.keys()):
        return False
    text = str(model)
    m = re.search(r": (.*?)Model", text) or re.search(r": (.*?)Stack", text) or re.match(r"(.*?)Model", text)
    return m is not None and m.group(1).lower() in ("bloom", "opt")


def replace_module(model, orig_class, replace_fn, _replace_policy, checkpoint=None):
    """Replace every instance of ``orig_class`` (or, with ``orig_class=None``, of every class a registered policy names) by
    ``replace_fn(child, policy, layer_id)``; ``checkpoint`` (``.pt`` / ``.safetensors``) is loaded once and the matching
    slice of its state dict is pushed into each replaced module (reference ``replace_module.py:619``)."""
    sd = None
    if checkpoint is not None:
        if str(checkpoint).endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(checkpoint)
        else:
            sd = torch.load(checkpoint, map_location="cpu", weights_only=False)
    policies = {}
    if orig_class is not None:
        policies[orig_class] = (replace_fn, _replace_policy)
    else:
        from .replace_policy import replace_policies
        for plcy in replace_policies:
            orig = getattr(plcy, "_orig_layer_class", None)
            for cls in (orig if isinstance(orig, (list, tuple)) else [orig]):
                if cls is not None:
                    policies[cls] = (replace_fn, plcy)
    assert policies, "No default policy found! Please specify your policy injection_policy (like {BertLayer:HFBEertLayerPolicy})."
    skip0 = skip_level_0_prefix(model, sd)
    counter = [0]

    def walk(mod, prefix, level):
        for name, child in list(mod.named_children()):
            full = prefix + name
            hit = policies.get(type(child))
            if hit is None:
                walk(child, full + "." if not (level == 0 and skip0) else prefix, level + 1)
                continue
            fn, pol = hit
            new = fn(child, pol, counter[0])
            counter[0] += 1
            if sd is not None:
                own = {k[len(full) + 1:]: v for k, v in sd.items() if k.startswith(full + ".")}
                if own:
                    new.load_state_dict(own, strict=False)
            setattr(mod, name, new)

    walk(model, "", 0)
    return model
