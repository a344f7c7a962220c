"""Architecture description shared by every decoder family the ragged engine serves.

The reference has one hand-written model + container + policy per family under
``inference/v2/model_implementations/{llama_v2,mistral,mixtral,falcon,opt,phi,phi3,qwen,qwen_v2,qwen_v2_moe}``.
Here the family differences are *data* (``ArchSpec``) consumed by one implementation
(``ragged_transformer.RaggedTransformer``) plus one weight-name map per family (``weights.py``).
"""
from dataclasses import dataclass, field
from typing import Optional


@dataclass
class ArchSpec:
    model_type: str
    vocab_size: int
    hidden: int
    layers: int
    heads: int
    kv_heads: int
    head_dim: int
    intermediate: int
    norm: str = "rms"                # rms | layer
    norm_eps: float = 1e-5
    act: str = "silu"                # silu | gelu | gelu_new | relu
    gated_mlp: bool = True
    qkv_bias: bool = False
    out_bias: bool = False
    mlp_bias: bool = False
    positional: str = "rope"         # rope | learned
    pos_offset: int = 0              # OPT stores positions shifted by 2
    rope_theta: float = 10000.0
    rope_scaling: Optional[dict] = None
    rotary_dim: Optional[int] = None     # partial rotary (phi)
    max_positions: int = 8192
    parallel_residual: bool = False  # falcon / phi-2: attn and mlp read the same normed input
    shared_ln: bool = False          # parallel residual with a single layernorm
    tie_embeddings: bool = False
    final_norm: bool = True
    lm_head_bias: bool = False
    # MoE
    num_experts: int = 0
    top_k: int = 0
    norm_topk: bool = True
    shared_expert_intermediate: int = 0
    extras: dict = field(default_factory=dict)

    @property
    def rot_dim(self):
        if self.positional != "rope":
            return 0
        return self.rotary_dim or self.head_dim


def _g(cfg, *names, default=None):
    for n in names:
        v = getattr(cfg, n, None) if not isinstance(cfg, dict) else cfg.get(n)
        if v is not None:
            return v
    return default


def _rope(cfg):
    rp = _g(cfg, "rope_parameters")
    theta = _g(cfg, "rope_theta")
    scaling = _g(cfg, "rope_scaling")
    if rp:
        theta = theta or rp.get("rope_theta")
        if rp.get("rope_type", "default") not in ("default", None):
            scaling = scaling or rp
    return float(theta or 10000.0), scaling


def _partial(cfg, *names, default=1.0):
    rp = _g(cfg, "rope_parameters") or {}
    v = _g(cfg, *names)
    if v is None:
        v = rp.get("partial_rotary_factor")
    return default if v is None else v


SUPPORTED_MODEL_TYPES = ("llama", "mistral", "mixtral", "qwen2", "qwen2_moe", "qwen", "phi3", "phi", "falcon", "opt",
                         "gpt2", "gpt_neox")


def arch_from_hf_config(cfg) -> ArchSpec:
    mt = _g(cfg, "model_type")
    if mt not in SUPPORTED_MODEL_TYPES:
        raise ValueError(f"Unsupported model type {mt}; supported: {SUPPORTED_MODEL_TYPES}")
    hidden = _g(cfg, "hidden_size", "n_embd", "d_model")
    heads = _g(cfg, "num_attention_heads", "n_head")
    layers = _g(cfg, "num_hidden_layers", "n_layer")
    kv = _g(cfg, "num_key_value_heads", default=heads)
    hd = _g(cfg, "head_dim", default=hidden // heads)
    theta, scaling = _rope(cfg)
    common = dict(model_type=mt, vocab_size=_g(cfg, "vocab_size"), hidden=hidden, layers=layers, heads=heads, kv_heads=kv,
                  head_dim=hd, intermediate=_g(cfg, "intermediate_size", "ffn_dim", "n_inner", default=4 * hidden),
                  max_positions=_g(cfg, "max_position_embeddings", "n_positions", default=8192), rope_theta=theta,
                  rope_scaling=scaling, tie_embeddings=bool(_g(cfg, "tie_word_embeddings", default=False)))
    if mt in ("llama", "mistral", "qwen2", "phi3"):
        return ArchSpec(**common, norm_eps=_g(cfg, "rms_norm_eps", default=1e-5), qkv_bias=(mt == "qwen2") or
                        bool(_g(cfg, "attention_bias", default=False)))
    if mt == "mixtral":
        return ArchSpec(**common, norm_eps=_g(cfg, "rms_norm_eps", default=1e-5), num_experts=_g(cfg, "num_local_experts"),
                        top_k=_g(cfg, "num_experts_per_tok"))
    if mt == "qwen2_moe":
        common["intermediate"] = _g(cfg, "moe_intermediate_size")
        return ArchSpec(**common, norm_eps=_g(cfg, "rms_norm_eps", default=1e-6), qkv_bias=True,
                        num_experts=_g(cfg, "num_experts"), top_k=_g(cfg, "num_experts_per_tok"),
                        norm_topk=bool(_g(cfg, "norm_topk_prob", default=False)),
                        shared_expert_intermediate=_g(cfg, "shared_expert_intermediate_size", default=0))
    if mt == "qwen":
        common["intermediate"] = _g(cfg, "intermediate_size") // 2
        return ArchSpec(**common, norm_eps=_g(cfg, "layer_norm_epsilon", default=1e-6), qkv_bias=True)
    if mt == "phi":
        return ArchSpec(**common, norm="layer", norm_eps=_g(cfg, "layer_norm_eps", default=1e-5), act="gelu_new",
                        gated_mlp=False, qkv_bias=True, out_bias=True, mlp_bias=True, parallel_residual=True,
                        shared_ln=True, rotary_dim=int(hd * _partial(cfg, "partial_rotary_factor", default=0.5)),
                        lm_head_bias=True)
    if mt == "falcon":
        new_arch = bool(_g(cfg, "new_decoder_architecture", default=False))
        kvh = _g(cfg, "num_kv_heads", default=heads) if (new_arch or not _g(cfg, "multi_query", default=True)) else 1
        common["kv_heads"] = kvh
        return ArchSpec(**common, norm="layer", norm_eps=_g(cfg, "layer_norm_epsilon", default=1e-5), act="gelu",
                        gated_mlp=False, qkv_bias=bool(_g(cfg, "bias", default=False)), parallel_residual=
                        bool(_g(cfg, "parallel_attn", default=True)), shared_ln=not new_arch,
                        extras={"new_decoder_architecture": new_arch})
    if mt == "opt":
        return ArchSpec(**common, norm="layer", act="relu", gated_mlp=False, qkv_bias=True, out_bias=True, mlp_bias=True,
                        positional="learned", pos_offset=2,
                        extras={"do_layer_norm_before": bool(_g(cfg, "do_layer_norm_before", default=True))})
    if mt == "gpt2":
        common["tie_embeddings"] = True
        return ArchSpec(**common, norm="layer", norm_eps=_g(cfg, "layer_norm_epsilon", default=1e-5), act="gelu_new",
                        gated_mlp=False, qkv_bias=True, out_bias=True, mlp_bias=True, positional="learned")
    if mt == "gpt_neox":
        return ArchSpec(**common, norm="layer", norm_eps=_g(cfg, "layer_norm_eps", default=1e-5), act="gelu",
                        gated_mlp=False, qkv_bias=True, out_bias=True, mlp_bias=True,
                        parallel_residual=bool(_g(cfg, "use_parallel_residual", default=True)),
                        rotary_dim=int(hd * _partial(cfg, "rotary_pct", default=0.25)))
    raise AssertionError(mt)
