"""Engine construction (reference ``inference/v2/engine_factory.py`` ``build_hf_engine :66``,
``build_engine_from_ds_checkpoint :22``) + HF checkpoint reader (``checkpoint/huggingface_engine.py``)."""
import json
import os
from typing import Optional

import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.accelerator import get_accelerator
from .config_v2 import RaggedInferenceEngineConfig
from .engine_v2 import InferenceEngineV2
from .model_implementations import ArchSpec, RaggedTransformer, arch_from_hf_config, load_hf_weights, weights_from_b200_model


def _tp(engine_config):
    tp = engine_config.tensor_parallel.tp_size
    if tp > 1:
        if not dist.is_initialized():
            dist.init_distributed()
        from deepspeed_b200.utils import groups
        if groups.ranks_of("tp") is None:
            groups._init_tp_mesh_device(tensor_model_parallel_size=tp)
        g = groups.get_tensor_model_parallel_group()
        return g, tp, dist.get_rank(g)
    return None, 1, 0


def _as_cfg(engine_config):
    if engine_config is None:
        return RaggedInferenceEngineConfig()
    if isinstance(engine_config, dict):
        return RaggedInferenceEngineConfig(**engine_config)
    return engine_config


class HuggingFaceCheckpointEngine:
    """Lazy reader over a local HF directory (safetensors or torch .bin shards); no hub access."""

    def __init__(self, path: str):
        self.path = path
        with open(os.path.join(path, "config.json")) as f:
            self.model_config = json.load(f)
        self._index = {}
        self._open = {}
        files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
        self._safe = bool(files)
        if not files:
            files = sorted(f for f in os.listdir(path) if f.endswith(".bin") or f.endswith(".pt"))
        for fn in files:
            full = os.path.join(path, fn)
            if self._safe:
                from safetensors import safe_open
                with safe_open(full, framework="pt") as f:
                    for k in f.keys():
                        self._index[k] = full
            else:
                sd = torch.load(full, map_location="cpu", weights_only=True)
                self._open[full] = sd
                for k in sd:
                    self._index[k] = full

    def get(self, name: str) -> Optional[torch.Tensor]:
        full = self._index.get(name)
        if full is None:
            return None
        if self._safe:
            from safetensors import safe_open
            with safe_open(full, framework="pt") as f:
                return f.get_tensor(name)
        return self._open[full][name]

    def parameters(self):
        for k in self._index:
            yield k, self.get(k)


def build_hf_engine(path, engine_config=None, debug_level=None, dtype=torch.bfloat16, device=None) -> InferenceEngineV2:
    """``path``: local HF checkpoint dir, or an in-memory HF ``PreTrainedModel``."""
    engine_config = _as_cfg(engine_config)
    group, tp, rank = _tp(engine_config)
    device = device or get_accelerator().current_device_name()
    if isinstance(path, str):
        ck = HuggingFaceCheckpointEngine(path)
        spec = arch_from_hf_config(ck.model_config)
        get = ck.get
    else:
        spec = arch_from_hf_config(path.config)
        sd = path.state_dict()
        get = sd.get
    model = RaggedTransformer(spec, tp_group=group, tp_size=tp, tp_rank=rank, dtype=dtype, device=device)
    load_hf_weights(model, get, engine_config.quantization.quantization_mode)
    return InferenceEngineV2(model, engine_config, tp_group=group)


def build_engine_from_model(module, engine_config=None, dtype=None, device=None, tp_override=None) -> InferenceEngineV2:
    """Serve one of this repo's training models (``deepspeed_b200.models``) directly.  ``tp_override`` =
    ``(group, size, rank)`` shards over an explicit process group (the hybrid engine's generation-time TP group)
    instead of the global tensor-parallel mesh."""
    engine_config = _as_cfg(engine_config)
    group, tp, rank = tp_override if tp_override is not None else _tp(engine_config)
    cfg = module.cfg
    p = next(module.parameters())
    dtype, device = dtype or p.dtype, device or p.device
    name = type(module).__name__.lower()
    if "gpt2" in name:
        spec = arch_from_hf_config(dict(model_type="gpt2", vocab_size=cfg.vocab_size, n_embd=cfg.n_embd, n_head=cfg.n_head,
                                        n_layer=cfg.n_layer, n_positions=cfg.n_positions,
                                        layer_norm_epsilon=cfg.layer_norm_epsilon))
    else:
        mt = "mixtral" if "mixtral" in name else "llama"
        d = dict(model_type=mt, vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, num_attention_heads=
                 cfg.num_attention_heads, num_hidden_layers=cfg.num_hidden_layers, num_key_value_heads=
                 cfg.num_key_value_heads, head_dim=cfg.head_dim, intermediate_size=cfg.intermediate_size,
                 max_position_embeddings=cfg.max_position_embeddings, rope_theta=cfg.rope_theta,
                 rope_scaling=getattr(cfg, "rope_scaling", None), rms_norm_eps=cfg.rms_norm_eps,
                 tie_word_embeddings=getattr(cfg, "tie_word_embeddings", False),
                 attention_bias=getattr(cfg, "attention_bias", False))
        if mt == "mixtral":
            d.update(num_local_experts=cfg.num_local_experts, num_experts_per_tok=cfg.num_experts_per_tok)
        spec = arch_from_hf_config(d)
    model = RaggedTransformer(spec, tp_group=group, tp_size=tp, tp_rank=rank, dtype=dtype, device=device)
    weights_from_b200_model(model, module, engine_config.quantization.quantization_mode)
    return InferenceEngineV2(model, engine_config, tp_group=group)


def build_engine_from_ds_checkpoint(path: str, engine_config=None, debug_level=None) -> InferenceEngineV2:
    """Re-create an engine from ``InferenceEngineV2.serialize(path)`` output: the per-rank, already sharded / fused weights
    are loaded as they are (no checkpoint re-mapping)."""
    import torch
    from .model_implementations.arch import ArchSpec
    from .model_implementations.ragged_transformer import LayerWeights, RaggedTransformer
    engine_config = _as_cfg(engine_config)
    from deepspeed_b200 import comm as dist
    tp = engine_config.tensor_parallel.tp_size
    rank = dist.get_rank() % tp if (dist.is_initialized() and tp > 1) else 0
    from .model_implementations.flat_model_helpers import make_param_filename
    f = make_param_filename(path, rank, tp)
    if not os.path.exists(f):
        f = os.path.join(path, f"params_rank_{rank}.pt")  # layout written by earlier versions
    blob = torch.load(f, map_location="cpu", weights_only=False)
    assert blob["tp_size"] == tp, f"checkpoint was serialized for tp_size={blob['tp_size']}, engine config asks for {tp}"
    spec = ArchSpec(**blob["spec"])
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sample = next(v for v in blob["globals"].values() if torch.is_tensor(v))
    group = None
    if tp > 1:
        from deepspeed_b200.utils import groups
        if groups.ranks_of("tp") is None:
            groups._init_tp_mesh_device(tensor_model_parallel_size=tp)
        group = groups.get_tensor_model_parallel_group()
    model = RaggedTransformer(spec, group, tp, rank, sample.dtype, dev)

    def put(v):
        if torch.is_tensor(v):
            return v.to(dev)
        if isinstance(v, list):
            return [put(x) for x in v]
        if hasattr(v, "q") and hasattr(v, "params"):
            v.q, v.params = v.q.to(dev), v.params.to(dev)
        return v

    for k, v in blob["globals"].items():
        setattr(model, k, put(v))
    for lw, saved in zip(model.layers, blob["layers"]):
        for s, v in saved.items():
            setattr(lw, s, put(v))
    return InferenceEngineV2(model, engine_config, tp_group=group)
