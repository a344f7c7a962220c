"""Building blocks for the checkpoint-name -> parameter-dependency tables of the family containers.

Every Hugging Face decoder names its tensors a little differently but the *shapes* of the mapping repeat: separate or
fused q/k/v, an output projection, a plain or gated MLP, one or two norms per layer, optional biases.  The family
containers compose these helpers instead of spelling the dictionaries out.
"""


def _wb(src, dst, dep, bias, bias_dst=None):
    out = {f"{src}.weight": f"{dst}.{dep}"}
    if bias:
        out[f"{src}.bias"] = f"{bias_dst or dst.replace('_w', '_b')}.{dep}"
    return out


def split_qkv(attn, q="q_proj", k="k_proj", v="v_proj", bias=False):
    """Three projections -> ``qkv_w`` (and ``qkv_b``) dependencies q_params / k_params / v_params."""
    out = {}
    for mod, dep in ((q, "q_params"), (k, "k_params"), (v, "v_params")):
        out.update(_wb(f"{attn}.{mod}", "qkv_w", dep, bias))
    return out


def fused_qkv(name, bias=False):
    return _wb(name, "qkv_w", "params", bias)


def attn_out(name, bias=False):
    return _wb(name, "attn_out_w", "params", bias)


def gated_mlp(gate, up, down):
    return {f"{gate}.weight": "mlp_1_w.gate_params", f"{up}.weight": "mlp_1_w.up_params", f"{down}.weight": "mlp_2_w.params"}


def plain_mlp(fc1, fc2, bias=False):
    return {**_wb(fc1, "mlp_1_w", "params", bias), **_wb(fc2, "mlp_2_w", "params", bias)}


def norm(name, dst, bias_dst=None):
    out = {f"{name}.weight": f"{dst}.params"}
    if bias_dst:
        out[f"{name}.bias"] = f"{bias_dst}.params"
    return out


def routed_experts(router, experts, gate, up, down):
    """Top-k MoE: router + per-expert gated MLPs (``*`` is the expert index)."""
    return {f"{router}.weight": "moe_gate.params", f"{experts}.*.{gate}.weight": "moe_mlp_1.gating_experts",
            f"{experts}.*.{up}.weight": "moe_mlp_1.up_experts", f"{experts}.*.{down}.weight": "moe_mlp_2.experts"}


def embeddings(embed, final_norm, lm_head=None, final_norm_bias=False, tie=False):
    """Non-transformer container: token embedding (+ tied LM head), final norm, LM head."""
    out = {f"{embed}.weight": ["word_emb.params", "word_unembed.params"] if tie else "word_emb.params"}
    out.update(norm(final_norm, "final_norm_w" if final_norm_bias else "final_norm", "final_norm_b" if final_norm_bias else None))
    if lm_head and not tie:
        out[f"{lm_head}.weight"] = "word_unembed.params"
    return out
