"""Serialise / restore a built inference model as ONE flat buffer + metadata (reference
``model_implementations/flat_model_helpers.py``): restart a server without re-reading, re-fusing, re-sharding and
re-quantising the original checkpoint."""
import json
import os
from typing import Dict, Tuple

import torch

from typing import Optional

from deepspeed_b200.runtime.config_utils import DeepSpeedConfigModel

ALIGN = 256  # every tensor starts on a 256-byte boundary (TMA / vector friendly)


class TensorMetadata(DeepSpeedConfigModel):
    """Where one tensor lives in the flat buffer."""
    dtype: Optional[str] = None
    shape: Optional[Tuple[int, ...]] = None
    strides: Optional[Tuple[int, ...]] = None
    offset: int


class ParameterMetadata(DeepSpeedConfigModel):
    """A parameter = its main tensor + auxiliary tensors (quantisation scales ...)."""
    core_param: Optional[TensorMetadata] = None
    aux_params: Dict[str, TensorMetadata] = {}


class LayerMetadata(DeepSpeedConfigModel):
    params: Dict[str, ParameterMetadata] = {}


class ModelMetadata(DeepSpeedConfigModel):
    """``layers``: ``"<i>"`` for transformer layers, ``"non_transformer"`` for embeddings / final norm / unembed."""
    policy: str = ""
    layers: Dict[str, LayerMetadata] = {}


def make_param_filename(base: str, rank: int, n_ranks: int) -> str:
    return os.path.join(base, f"params_rank_{rank}_of_{n_ranks}.pt")


def make_metadata_filename(base: str, rank: int, n_ranks: int) -> str:
    return os.path.join(base, f"metadata_rank_{rank}_of_{n_ranks}.json")


def make_model_config_filename(base: str) -> str:
    return os.path.join(base, "ds_model_config.json")


def to_model_metadata(meta: dict, policy: str = "") -> ModelMetadata:
    """Group the flat ``name -> record`` table by layer / parameter (``x.scales`` is an aux tensor of ``x.q``)."""
    layers: Dict[str, dict] = {}
    for name, m in meta.items():
        parts = name.split(".")
        if parts[0] == "layers" and len(parts) > 2:
            layer, pname = parts[1], ".".join(parts[2:])
        else:
            layer, pname = "non_transformer", name
        strides, acc = [], 1
        for d in reversed(m["shape"]):
            strides.insert(0, acc)
            acc *= d
        tm = TensorMetadata(dtype=m["dtype"], shape=tuple(m["shape"]), strides=tuple(strides), offset=m["offset"])
        params = layers.setdefault(layer, {})
        if pname.endswith(".scales"):
            params.setdefault(pname[:-7], {"core": None, "aux": {}})["aux"]["scales"] = tm
        else:
            key = pname[:-2] if pname.endswith(".q") else pname
            params.setdefault(key, {"core": None, "aux": {}})["core"] = tm
    return ModelMetadata(policy=policy, layers={
        l: LayerMetadata(params={k: ParameterMetadata(core_param=v["core"], aux_params=v["aux"]) for k, v in ps.items()})
        for l, ps in layers.items()})


def pad_to_aligned_offset(offset: int, alignment: int = ALIGN) -> int:
    return -(-offset // alignment) * alignment


def _named_tensors(model) -> Dict[str, torch.Tensor]:
    out = {}
    if isinstance(model, torch.nn.Module):
        for n, p in model.named_parameters():
            out[n] = p.data
        for n, b in model.named_buffers():
            out["buffer:" + n] = b
    extra = getattr(model, "flat_tensors", None)  # models keeping weights outside nn.Parameter expose them here
    if callable(extra):
        out.update(extra())
    return out


def flatten_inference_model(model, path_prefix: str = None) -> Tuple[torch.Tensor, dict]:
    """-> (uint8 buffer, metadata).  With ``path_prefix``: also writes ``<prefix>.bin`` and ``<prefix>.json``."""
    tensors = _named_tensors(model)
    meta, offset = {}, 0
    for name, t in tensors.items():
        offset = pad_to_aligned_offset(offset)
        nbytes = t.numel() * t.element_size()
        meta[name] = {"offset": offset, "shape": list(t.shape), "dtype": str(t.dtype), "nbytes": nbytes}
        offset += nbytes
    buf = torch.zeros(pad_to_aligned_offset(offset), dtype=torch.uint8)
    for name, t in tensors.items():
        m = meta[name]
        buf[m["offset"]:m["offset"] + m["nbytes"]].copy_(t.detach().contiguous().cpu().view(-1).view(torch.uint8))
    if path_prefix is not None:
        os.makedirs(os.path.dirname(os.path.abspath(path_prefix)), exist_ok=True)
        buf.numpy().tofile(path_prefix + ".bin")
        with open(path_prefix + ".json", "w") as f:
            json.dump(meta, f)
    return buf, meta


def restore_inference_model(model, path_prefix_or_buffer, metadata: dict = None) -> None:
    """Copy every tensor back from the flat buffer (file prefix or in-memory buffer + metadata)."""
    if isinstance(path_prefix_or_buffer, str):
        import numpy as np
        with open(path_prefix_or_buffer + ".json") as f:
            metadata = json.load(f)
        buf = torch.from_numpy(np.fromfile(path_prefix_or_buffer + ".bin", dtype=np.uint8))
    else:
        buf = path_prefix_or_buffer
    from deepspeed_b200.inference.v2.inference_parameter import STR_TO_DTYPE
    tensors = _named_tensors(model)
    missing = [n for n in tensors if n not in metadata]
    if missing:
        raise KeyError(f"flat model is missing {missing[:5]}")
    with torch.no_grad():
        for name, t in tensors.items():
            m = metadata[name]
            if list(t.shape) != m["shape"] or str(t.dtype) != m["dtype"]:
                raise ValueError(f"{name}: flat model has {m['shape']} {m['dtype']}, model expects {list(t.shape)} {t.dtype}")
            src = buf[m["offset"]:m["offset"] + m["nbytes"]].view(STR_TO_DTYPE[m["dtype"]]).view(t.shape)
            t.copy_(src)
