"""Checkpoint-driven weight loading for injected models (reference ``module_inject/load_checkpoint.py``).

After kernel injection the model contains ``InjectedLayer`` wrappers (fused layers) plus untouched modules (embeddings,
final norm, heads).  ``load_model_with_checkpoint`` walks the model with the ORIGINAL parameter names of the checkpoint:
untouched modules load by name (tensor-parallel linears sliced on the fly), fused layers load through their family's
``param_names`` mapping (q/k/v concatenation, gate/up stacking, per-head de-interleave).
"""
import torch
from torch import nn

from deepspeed_b200.utils import logger

from .containers.base import InjectedLayer
from .policy import deinterleave_qkv

# checkpoint parameter suffixes -> fused-layer parameter, per family signature (first matching family wins)
_FAMILIES = [
    {  # llama / mistral / qwen2 / internlm
        "probe": "self_attn.q_proj.weight", "probe2": "mlp.gate_proj.weight",
        "qkvw": ["self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight"],
        "qkvb": ["self_attn.q_proj.bias", "self_attn.k_proj.bias", "self_attn.v_proj.bias"],
        "ow": "self_attn.o_proj.weight", "ob": "self_attn.o_proj.bias",
        "w1": ["mlp.gate_proj.weight", "mlp.up_proj.weight"], "b1": None, "w2": "mlp.down_proj.weight", "b2": None,
        "in_nw": "input_layernorm.weight", "in_nb": None, "at_nw": "post_attention_layernorm.weight", "at_nb": None},
    {  # opt
        "probe": "self_attn.q_proj.weight", "probe2": "fc1.weight",
        "qkvw": ["self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight"],
        "qkvb": ["self_attn.q_proj.bias", "self_attn.k_proj.bias", "self_attn.v_proj.bias"],
        "ow": "self_attn.out_proj.weight", "ob": "self_attn.out_proj.bias", "w1": "fc1.weight", "b1": "fc1.bias",
        "w2": "fc2.weight", "b2": "fc2.bias", "in_nw": "self_attn_layer_norm.weight", "in_nb": "self_attn_layer_norm.bias",
        "at_nw": "final_layer_norm.weight", "at_nb": "final_layer_norm.bias"},
    {  # gpt2 (Conv1D: transposed)
        "probe": "attn.c_attn.weight", "probe2": "mlp.c_fc.weight", "transposed": True,
        "qkvw": "attn.c_attn.weight", "qkvb": "attn.c_attn.bias", "ow": "attn.c_proj.weight", "ob": "attn.c_proj.bias",
        "w1": "mlp.c_fc.weight", "b1": "mlp.c_fc.bias", "w2": "mlp.c_proj.weight", "b2": "mlp.c_proj.bias",
        "in_nw": "ln_1.weight", "in_nb": "ln_1.bias", "at_nw": "ln_2.weight", "at_nb": "ln_2.bias"},
    {  # bloom / gpt-neox (per-head fused qkv)
        "probe": "self_attention.query_key_value.weight", "probe2": "mlp.dense_h_to_4h.weight", "interleaved": True,
        "qkvw": "self_attention.query_key_value.weight", "qkvb": "self_attention.query_key_value.bias",
        "ow": "self_attention.dense.weight", "ob": "self_attention.dense.bias", "w1": "mlp.dense_h_to_4h.weight",
        "b1": "mlp.dense_h_to_4h.bias", "w2": "mlp.dense_4h_to_h.weight", "b2": "mlp.dense_4h_to_h.bias",
        "in_nw": "input_layernorm.weight", "in_nb": "input_layernorm.bias", "at_nw": "post_attention_layernorm.weight",
        "at_nb": "post_attention_layernorm.bias"},
    {
        "probe": "attention.query_key_value.weight", "probe2": "mlp.dense_h_to_4h.weight", "interleaved": True,
        "qkvw": "attention.query_key_value.weight", "qkvb": "attention.query_key_value.bias", "ow": "attention.dense.weight",
        "ob": "attention.dense.bias", "w1": "mlp.dense_h_to_4h.weight", "b1": "mlp.dense_h_to_4h.bias",
        "w2": "mlp.dense_4h_to_h.weight", "b2": "mlp.dense_4h_to_h.bias", "in_nw": "input_layernorm.weight",
        "in_nb": "input_layernorm.bias", "at_nw": "post_attention_layernorm.weight", "at_nb": "post_attention_layernorm.bias"},
    {  # bert
        "probe": "attention.self.query.weight", "probe2": "intermediate.dense.weight",
        "qkvw": ["attention.self.query.weight", "attention.self.key.weight", "attention.self.value.weight"],
        "qkvb": ["attention.self.query.bias", "attention.self.key.bias", "attention.self.value.bias"],
        "ow": "attention.output.dense.weight", "ob": "attention.output.dense.bias", "w1": "intermediate.dense.weight",
        "b1": "intermediate.dense.bias", "w2": "output.dense.weight", "b2": "output.dense.bias",
        "in_nw": "attention.output.LayerNorm.weight", "in_nb": "attention.output.LayerNorm.bias",
        "at_nw": "output.LayerNorm.weight", "at_nb": "output.LayerNorm.bias"},
]


def _get(sd, prefix, names, transposed=False):
    if names is None:
        return None
    if isinstance(names, (list, tuple)):
        parts = [sd.get(prefix + n) for n in names]
        return None if any(p is None for p in parts) else torch.cat(parts, dim=0)
    t = sd.get(prefix + names)
    return t.t() if (t is not None and transposed and t.dim() == 2) else t


def _tp_slice_fused(fused, tensors, tp, rank):
    """Apply the same slicing ``BaseTransformerContainer.apply_tensor_parallelism`` does, to checkpoint tensors."""
    if tp <= 1:
        return tensors
    c = fused.config
    heads, kv = c.heads, (c.num_kv if c.num_kv > 0 else c.heads)
    d = c.hidden_size // heads
    out = dict(tensors)

    def split_qkv(t):
        if t is None:
            return None
        q, k, v = t[:heads * d], t[heads * d:(heads + kv) * d], t[(heads + kv) * d:]
        take = lambda x, n: x[rank * (n // tp) * d:(rank + 1) * (n // tp) * d]
        return torch.cat([take(q, heads), take(k, kv), take(v, kv)], 0)

    gated = tensors["w1"] is not None and tensors["w1"].shape[0] == 2 * c.intermediate_size
    rows = lambda t, parts: None if t is None else torch.cat([x.chunk(tp, 0)[rank] for x in t.chunk(parts, 0)], 0)
    out["qkvw"], out["qkvb"] = split_qkv(tensors["qkvw"]), split_qkv(tensors["qkvb"])
    out["ow"] = tensors["ow"].chunk(tp, 1)[rank]
    out["w1"], out["b1"] = rows(tensors["w1"], 2 if gated else 1), rows(tensors["b1"], 2 if gated else 1)
    out["w2"] = tensors["w2"].chunk(tp, 1)[rank]
    if rank != 0:
        out["ob"] = None if tensors["ob"] is None else torch.zeros_like(tensors["ob"])
        out["b2"] = None if tensors["b2"] is None else torch.zeros_like(tensors["b2"])
    return out


def _load_fused(wrapper, sd, prefix, tp, rank):
    fused = wrapper.fused
    fam = next((f for f in _FAMILIES if prefix + f["probe"] in sd and prefix + f["probe2"] in sd), None)
    if fam is None:
        return False
    tr = fam.get("transposed", False)
    t = {k: _get(sd, prefix, fam[k], tr) for k in ("qkvw", "qkvb", "ow", "ob", "w1", "b1", "w2", "b2", "in_nw", "in_nb", "at_nw",
                                                   "at_nb")}
    if fam.get("interleaved"):
        t["qkvw"], t["qkvb"] = deinterleave_qkv(t["qkvw"], fused.config.heads), deinterleave_qkv(t["qkvb"], fused.config.heads)
    t = _tp_slice_fused(fused, t, tp, rank)
    dst = {"qkvw": fused.attn_qkvw, "qkvb": fused.attn_qkvb, "ow": fused.attn_ow, "ob": fused.attn_ob, "w1": fused.inter_w,
           "b1": fused.inter_b, "w2": fused.output_w, "b2": fused.output_b, "in_nw": fused.norm_w, "in_nb": fused.norm_b,
           "at_nw": fused.attn_nw, "at_nb": fused.attn_nb}
    with torch.no_grad():
        for k, d in dst.items():
            if t[k] is None:
                d.zero_()
            else:
                d.copy_(t[k].to(d.dtype).reshape(d.shape))
    return True


def load_model_with_checkpoint(r_module, sd, mp_replace=None, ckpt_type="pp", ckpt_mp_size=1, weight_quantizer=None, rank=0,
                               container=None, mp_group=None, mp_size=1):
    """Load ``sd`` (original parameter names; a dict or a list of dicts to merge) into an injected model."""
    if isinstance(sd, (list, tuple)):
        merged = {}
        for part in sd:
            merged.update(part)
        sd = merged
    if mp_group is not None and mp_size > 1:
        from deepspeed_b200 import comm as dist
        rank = dist.get_rank(mp_group)
    loaded, missing = 0, []

    def walk(module, prefix):
        nonlocal loaded
        for name, child in module.named_children():
            p = f"{prefix}{name}."
            if isinstance(child, InjectedLayer):
                if _load_fused(child, sd, p, mp_size, rank):
                    loaded += 1
                else:
                    missing.append(p)
                continue
            for pn, param in child.named_parameters(recurse=False):
                t = sd.get(p + pn)
                if t is None:
                    continue
                with torch.no_grad():
                    if t.shape == param.shape:
                        param.copy_(t.to(param.dtype))
                    else:  # tensor-parallel linear / embedding: slice the dimension that differs
                        dim = next(i for i, (a, b) in enumerate(zip(t.shape, param.shape)) if a != b)
                        param.copy_(t.chunk(mp_size, dim)[rank].to(param.dtype))
                loaded += 1
            for bn, buf in child.named_buffers(recurse=False):
                if p + bn in sd and sd[p + bn].shape == buf.shape:
                    buf.copy_(sd[p + bn])
            walk(child, p)

    walk(r_module, "")
    # tied heads (lm_head <- embedding) when the checkpoint omits one of them
    for name, mod in r_module.named_modules():
        if name.endswith("lm_head") and isinstance(mod, nn.Linear) and f"{name}.weight" not in sd:
            emb = getattr(r_module, "get_input_embeddings", lambda: None)()
            if emb is not None and emb.weight.shape == mod.weight.shape:
                mod.weight = emb.weight
    if missing:
        logger.warning(f"load_model_with_checkpoint: no weights found for fused layers at {missing[:4]}...")
    return loaded
