"""Checkpoint-name maps + TP sharding for the ragged engine.

Role parity: the reference's per-family ``container.py`` ``PARAM_MAPPING`` tables and
``model_implementations/sharding/{qkv,mlp,attn_out,embedding,unembed}.py``.  Every family funnels into the
same packed layout: ``qkv_w`` = [q | k | v] rows for *this rank's* heads, ``up_w`` = [gate | up] rows, etc.
"""
import re
from typing import Callable, Dict, Optional

import torch

from .arch import ArchSpec
from .ragged_transformer import RaggedTransformer


def _rows(t, rank, world):
    n = t.shape[0] // world
    return t[rank * n:(rank + 1) * n]


def _cols(t, rank, world):
    n = t.shape[1] // world
    return t[:, rank * n:(rank + 1) * n]


def _shard_qkv(q, k, v, spec: ArchSpec, rank, world):
    """q [hq*d, H], k/v [hkv*d, H]; KV heads are replicated when hkv < world."""
    q = _rows(q, rank, world)
    if spec.kv_heads >= world:
        k, v = _rows(k, rank, world), _rows(v, rank, world)
    else:
        per = world // spec.kv_heads
        d = spec.head_dim
        h = rank // per
        k, v = k[h * d:(h + 1) * d], v[h * d:(h + 1) * d]
    return torch.cat([q, k, v], dim=0)


def _finish(model: RaggedTransformer, quant_mode=None):
    from deepspeed_b200.inference.quantization.layers import quantize_weight

    def put(t):
        if t is None:
            return None
        return t.to(device=model.device, dtype=model.dtype).contiguous()

    for lw in model.layers:
        for s in lw.__slots__:
            v = getattr(lw, s)
            if isinstance(v, list):
                if s in ("experts_up", "experts_down") and v and not quant_mode and len({tuple(x.shape) for x in v}) == 1:
                    # stacked [E, N, K]: the layout the grouped (MoE) GEMM kernel consumes; still indexable per expert.
                    # Filled expert by expert so the host never holds a second full copy.
                    stacked = torch.empty(len(v), *v[0].shape, dtype=model.dtype, device=model.device)
                    for i, x in enumerate(v):
                        stacked[i].copy_(x)
                    setattr(lw, s, stacked)
                else:
                    setattr(lw, s, [put(x) for x in v])
            elif v is not None:
                setattr(lw, s, put(v))
        if quant_mode:
            for s in ("qkv_w", "o_w", "up_w", "down_w"):
                if getattr(lw, s) is not None:
                    setattr(lw, s, quantize_weight(getattr(lw, s), quant_mode))
    for s in ("embed_w", "pos_w", "final_ln_w", "final_ln_b", "lm_head_w", "lm_head_b"):
        setattr(model, s, put(getattr(model, s)))
    return model


def _split_fused_qkv(w, spec: ArchSpec, layout):
    """Return (q, k, v) from a fused tensor.  layout: 'qkv' contiguous blocks, 'interleaved' per-head [q,k,v]
    (gpt-neox, falcon-old multi-head), 'grouped' falcon new-arch [group: q.. k v]."""
    hq, hkv, d = spec.heads, spec.kv_heads, spec.head_dim
    rest = w.shape[1:]
    if layout == "qkv":
        return torch.split(w, [hq * d, hkv * d, hkv * d], dim=0)
    if layout == "interleaved":
        x = w.view(hq, 3, d, *rest)
        return tuple(x[:, i].reshape(hq * d, *rest) for i in range(3))
    if layout == "grouped":
        per = hq // hkv
        x = w.view(hkv, per + 2, d, *rest)
        return (x[:, :per].reshape(hq * d, *rest), x[:, per].reshape(hkv * d, *rest), x[:, per + 1].reshape(hkv * d, *rest))
    raise ValueError(layout)


# per-family: prefix of layer i, and a dict of logical-name -> HF sub-name(s)
def _family_map(mt, spec):
    if mt in ("llama", "mistral", "qwen2", "mixtral", "qwen2_moe"):
        m = dict(layers="model.layers.{}", embed="model.embed_tokens.weight", final_ln="model.norm", lm_head="lm_head",
                 ln1="input_layernorm", ln2="post_attention_layernorm", q="self_attn.q_proj", k="self_attn.k_proj",
                 v="self_attn.v_proj", o="self_attn.o_proj", gate="mlp.gate_proj", up="mlp.up_proj", down="mlp.down_proj")
        if mt == "mixtral":
            m.update(router="block_sparse_moe.gate", expert="block_sparse_moe.experts.{}", e_gate="w1", e_up="w3", e_down="w2")
        if mt == "qwen2_moe":
            m.update(router="mlp.gate", expert="mlp.experts.{}", e_gate="gate_proj", e_up="up_proj", e_down="down_proj",
                     shared="mlp.shared_expert", shared_gate="mlp.shared_expert_gate")
        return m
    if mt == "phi3":
        return dict(layers="model.layers.{}", embed="model.embed_tokens.weight", final_ln="model.norm", lm_head="lm_head",
                    ln1="input_layernorm", ln2="post_attention_layernorm", qkv="self_attn.qkv_proj", qkv_layout="qkv",
                    o="self_attn.o_proj", gate_up="mlp.gate_up_proj", down="mlp.down_proj")
    if mt == "phi":
        return dict(layers="model.layers.{}", embed="model.embed_tokens.weight", final_ln="model.final_layernorm",
                    lm_head="lm_head", ln1="input_layernorm", q="self_attn.q_proj", k="self_attn.k_proj",
                    v="self_attn.v_proj", o="self_attn.dense", up="mlp.fc1", down="mlp.fc2")
    if mt == "falcon":
        new = spec.extras.get("new_decoder_architecture")
        return dict(layers="transformer.h.{}", embed="transformer.word_embeddings.weight", final_ln="transformer.ln_f",
                    lm_head="lm_head", ln1="ln_attn" if new else "input_layernorm", ln2="ln_mlp" if new else None,
                    qkv="self_attention.query_key_value", qkv_layout="grouped" if (new or spec.kv_heads == 1) else
                    "interleaved", o="self_attention.dense", up="mlp.dense_h_to_4h", down="mlp.dense_4h_to_h")
    if mt == "opt":
        return dict(layers="model.decoder.layers.{}", embed="model.decoder.embed_tokens.weight",
                    pos="model.decoder.embed_positions.weight", final_ln="model.decoder.final_layer_norm",
                    lm_head="lm_head", ln1="self_attn_layer_norm", ln2="final_layer_norm", q="self_attn.q_proj",
                    k="self_attn.k_proj", v="self_attn.v_proj", o="self_attn.out_proj", up="fc1", down="fc2")
    if mt == "gpt2":
        return dict(layers="transformer.h.{}", embed="transformer.wte.weight", pos="transformer.wpe.weight",
                    final_ln="transformer.ln_f", lm_head="lm_head", ln1="ln_1", ln2="ln_2", qkv="attn.c_attn",
                    qkv_layout="qkv", o="attn.c_proj", up="mlp.c_fc", down="mlp.c_proj", conv1d=True)
    if mt == "gpt_neox":
        return dict(layers="gpt_neox.layers.{}", embed="gpt_neox.embed_in.weight", final_ln="gpt_neox.final_layer_norm",
                    lm_head="embed_out", ln1="input_layernorm", ln2="post_attention_layernorm",
                    qkv="attention.query_key_value", qkv_layout="interleaved", o="attention.dense",
                    up="mlp.dense_h_to_4h", down="mlp.dense_4h_to_h")
    if mt == "qwen":
        return dict(layers="transformer.h.{}", embed="transformer.wte.weight", final_ln="transformer.ln_f",
                    lm_head="lm_head", ln1="ln_1", ln2="ln_2", qkv="attn.c_attn", qkv_layout="qkv", o="attn.c_proj",
                    gate="mlp.w2", up="mlp.w1", down="mlp.c_proj")
    raise ValueError(mt)


def load_hf_weights(model: RaggedTransformer, get: Callable[[str], Optional[torch.Tensor]], quant_mode=None):
    """``get(name)`` returns the full (unsharded) checkpoint tensor or None."""
    sp, r, w = model.spec, model.tp_rank, model.tp_size
    fm = _family_map(sp.model_type, sp)
    conv1d = fm.get("conv1d", False)

    def W(name):
        t = get(name + ".weight")
        if t is not None and conv1d and t.dim() == 2 and name != fm["lm_head"]:
            t = t.t()
        return t

    def B(name):
        return get(name + ".bias")

    model.embed_w = get(fm["embed"])
    if fm.get("pos"):
        model.pos_w = get(fm["pos"])
    if sp.final_norm:
        model.final_ln_w, model.final_ln_b = get(fm["final_ln"] + ".weight"), B(fm["final_ln"])
    head = W(fm["lm_head"])
    model.lm_head_w = head if head is not None else model.embed_w
    model.lm_head_b = B(fm["lm_head"]) if sp.lm_head_bias else None
    for i, lw in enumerate(model.layers):
        p = fm["layers"].format(i) + "."
        lw.ln1_w, lw.ln1_b = get(p + fm["ln1"] + ".weight"), B(p + fm["ln1"])
        if fm.get("ln2"):
            lw.ln2_w, lw.ln2_b = get(p + fm["ln2"] + ".weight"), B(p + fm["ln2"])
        if "qkv" in fm:
            q, k, v = _split_fused_qkv(W(p + fm["qkv"]), sp, fm["qkv_layout"])
            qb = B(p + fm["qkv"])
            bq = _split_fused_qkv(qb, sp, fm["qkv_layout"]) if qb is not None else None
        else:
            q, k, v = W(p + fm["q"]), W(p + fm["k"]), W(p + fm["v"])
            bq = (B(p + fm["q"]), B(p + fm["k"]), B(p + fm["v"])) if B(p + fm["q"]) is not None else None
        lw.qkv_w = _shard_qkv(q, k, v, sp, r, w)
        if bq is not None:
            lw.qkv_b = _shard_qkv(bq[0][:, None], bq[1][:, None], bq[2][:, None], sp, r, w).squeeze(1)
        lw.o_w, lw.o_b = _cols(W(p + fm["o"]), r, w), B(p + fm["o"])
        if sp.num_experts:
            lw.gate_w = W(p + fm["router"])
            fused_gu = fused_dn = None
            if lw.gate_w is None:  # transformers>=5 layout: mlp.gate + fused mlp.experts.{gate_up_proj,down_proj}
                lw.gate_w = W(p + "mlp.gate")
            if W(p + fm["expert"].format(0) + "." + fm["e_gate"]) is None:
                fused_gu, fused_dn = get(p + "mlp.experts.gate_up_proj"), get(p + "mlp.experts.down_proj")
            lw.experts_up, lw.experts_down = [], []
            for e in range(sp.num_experts):
                if fused_gu is not None:
                    g, u = fused_gu[e].chunk(2, dim=0)
                    dn = fused_dn[e]
                else:
                    ep = p + fm["expert"].format(e) + "."
                    g, u, dn = W(ep + fm["e_gate"]), W(ep + fm["e_up"]), W(ep + fm["e_down"])
                lw.experts_up.append(torch.cat([_rows(g, r, w), _rows(u, r, w)], 0))
                lw.experts_down.append(_cols(dn, r, w))
            if fm.get("shared") and W(p + fm["shared"] + ".gate_proj") is not None:
                s = p + fm["shared"]
                lw.shared_up = torch.cat([_rows(W(s + ".gate_proj"), r, w), _rows(W(s + ".up_proj"), r, w)], 0)
                lw.shared_down = _cols(W(s + ".down_proj"), r, w)
                lw.shared_gate = W(p + fm["shared_gate"])
        elif "gate_up" in fm:
            g, u = W(p + fm["gate_up"]).chunk(2, dim=0)
            lw.up_w = torch.cat([_rows(g, r, w), _rows(u, r, w)], 0)
            lw.down_w = _cols(W(p + fm["down"]), r, w)
        elif sp.gated_mlp:
            lw.up_w = torch.cat([_rows(W(p + fm["gate"]), r, w), _rows(W(p + fm["up"]), r, w)], 0)
            lw.down_w = _cols(W(p + fm["down"]), r, w)
        else:
            lw.up_w, lw.up_b = _rows(W(p + fm["up"]), r, w), (_rows(B(p + fm["up"]), r, w) if B(p + fm["up"]) is not None else None)
            lw.down_w, lw.down_b = _cols(W(p + fm["down"]), r, w), B(p + fm["down"])
    return _finish(model, quant_mode)


def weights_from_b200_model(model: RaggedTransformer, module, quant_mode=None):
    """Build from this repo's own training models (``models/{llama,mixtral,gpt2}.py``), whose layout is already
    packed; only TP slicing is needed."""
    sp, r, w = model.spec, model.tp_rank, model.tp_size
    sd = {k: v.detach() for k, v in module.state_dict().items()}
    if sp.model_type == "gpt2":
        def get(name):
            m = {"transformer.wte.weight": "wte.weight", "transformer.wpe.weight": "wpe.weight"}
            name = m.get(name, name).replace("transformer.", "")
            name = name.replace(".attn.c_attn", ".c_attn").replace(".attn.c_proj", ".c_proj")
            name = name.replace(".mlp.c_fc", ".c_fc").replace(".mlp.c_proj", ".c_proj2")
            t = sd.get(name)
            # this repo's GPT-2 uses nn.Linear ([out,in]); the HF map expects Conv1D ([in,out]) for 2-D weights
            if t is not None and t.dim() == 2 and re.search(r"(c_attn|c_proj2?|c_fc)\.weight$", name):
                t = t.t()
            return t
        return load_hf_weights(model, get, quant_mode)
    # llama / mixtral families: packed names -> unpack views to reuse the generic loader
    def get(name):
        if name in sd:
            return sd[name]
        m = re.match(r"(.*)\.self_attn\.([qkv])_proj\.(weight|bias)$", name)
        if m:
            t = sd.get(f"{m.group(1)}.self_attn.qkv_proj.{m.group(3)}")
            if t is None:
                return None
            q, k, v = torch.split(t, [sp.heads * sp.head_dim, sp.kv_heads * sp.head_dim, sp.kv_heads * sp.head_dim], 0)
            return {"q": q, "k": k, "v": v}[m.group(2)]
        m = re.match(r"(.*)\.mlp\.(gate|up)_proj\.weight$", name)
        if m:
            t = sd.get(f"{m.group(1)}.mlp.gate_up_proj.weight")
            return None if t is None else t.chunk(2, 0)[0 if m.group(2) == "gate" else 1]
        m = re.match(r"(.*)\.block_sparse_moe\.experts\.(\d+)\.(w1|w2|w3)\.weight$", name)
        if m:
            e = int(m.group(2))
            base = m.group(1) + ".block_sparse_moe."
            w13 = next((sd[k] for k in sd if k.startswith(base) and k.endswith("w13")), None)
            w2 = next((sd[k] for k in sd if k.startswith(base) and k.endswith("w2")), None)
            if w13 is None:
                return None
            return {"w1": w13[e].chunk(2, 0)[0], "w3": w13[e].chunk(2, 0)[1], "w2": w2[e]}[m.group(3)]
        if name.endswith("block_sparse_moe.gate.weight"):
            base = name[:-len("gate.weight")]
            return next((sd[k] for k in sd if k.startswith(base) and k.endswith("wg.weight")), None)
        return None
    return load_hf_weights(model, get, quant_mode)
