"""Policies tie a checkpoint format to a model implementation (reference ``model_implementations/inference_policy_base.py``)."""
from abc import ABC, ABCMeta, abstractmethod
from typing import Any, Iterable, Optional, Tuple

import torch

from .arch import ArchSpec, arch_from_hf_config
from .ragged_transformer import RaggedTransformer
from .weights import load_hf_weights

POLICIES = {}  # model_type -> policy class
POLICIES_BY_NAME = {}  # class name -> policy class


class ContainerMap:
    """Where each checkpoint prefix goes: per-layer transformer containers, the non-transformer container, and prefixes to
    ignore (reference ``ContainerMap``)."""

    def __init__(self) -> None:
        self._transformer = (None, None)
        self._non_transformer = None
        self._unmapped = []

    def set_transformer_params(self, prefixes, containers) -> None:
        self._transformer = (prefixes if isinstance(prefixes, (list, tuple)) else [prefixes], containers)

    def set_non_transformer_params(self, container) -> None:
        self._non_transformer = container

    def set_unmapped_params(self, prefixes) -> None:
        self._unmapped = list(prefixes) if isinstance(prefixes, (list, tuple)) else [prefixes]

    @property
    def transformer_params(self):
        return self._transformer[1]

    @property
    def non_transformer_params(self):
        return self._non_transformer

    def map_param(self, name: str, tensor: torch.Tensor) -> bool:
        if any(name.startswith(u) for u in self._unmapped):
            return True
        prefixes, containers = self._transformer
        for pre in prefixes or []:
            if name.startswith(pre + "."):
                idx, _, rest = name[len(pre) + 1:].partition(".")
                if idx.isdigit() and containers is not None and int(idx) < len(containers):
                    return containers[int(idx)].set_dependency(rest, tensor)
        return self._non_transformer.set_dependency(name, tensor) if self._non_transformer is not None else False

    def validate(self) -> None:
        cs = list(self._transformer[1] or []) + ([self._non_transformer] if self._non_transformer else [])
        missing = [i for i, c in enumerate(cs) if not c.is_initialized]
        if missing:
            raise RuntimeError(f"containers {missing} are missing parameters after loading the checkpoint")


class PolicyMeta(ABCMeta):
    """Registers every concrete policy under its ``model_type`` (and its class name) in ``POLICIES`` at class creation
    (reference ``inference_policy_base.py:95``)."""

    def __new__(mcs, name, bases, dct):
        cls = super().__new__(mcs, name, bases, dct)
        if name != "InferenceV2Policy":
            POLICIES_BY_NAME[name] = cls
            if dct.get("model_type") or getattr(cls, "model_type", None):
                POLICIES[cls.model_type] = cls
        return cls


class InferenceV2Policy(ABC, metaclass=PolicyMeta):
    """``model_config``: the HF config (object or dict); ``checkpoint_engine``: yields ``(name, tensor)``."""
    model_type: str = None

    def __init__(self, model_config: Any, checkpoint_engine: Optional[Any] = None, inf_checkpoint_path: Optional[str] = None) -> None:
        self._model_config = model_config
        self._checkpoint_engine = checkpoint_engine
        self._inf_checkpoint_path = inf_checkpoint_path

    def arch_spec(self) -> ArchSpec:
        return arch_from_hf_config(self._model_config)

    @abstractmethod
    def instantiate_model(self, engine_config, mp_group=None) -> RaggedTransformer:
        ...

    def build_container_map(self):
        """Optional declarative view (containers); the default loader uses the family weight map directly."""
        return None

    def build_model(self, engine_config, mp_group=None) -> RaggedTransformer:
        """Instantiate the model and populate it from the checkpoint engine (or a serialized flat model)."""
        model = self.instantiate_model(engine_config, mp_group)
        if self._inf_checkpoint_path is not None:
            from .flat_model_helpers import restore_inference_model
            restore_inference_model(model, self._inf_checkpoint_path)
            return model
        if self._checkpoint_engine is not None:
            eng = self._checkpoint_engine
            getter = eng.get if hasattr(eng, "get") else dict(eng.parameters()).get
            qm = getattr(getattr(engine_config, "quantization", None), "quantization_mode", None)
            load_hf_weights(model, getter, qm)
        return model


def policy_for(model_type: str):
    from . import llama_v2, mistral, mixtral, opt, falcon, phi, phi3, qwen, qwen_v2, qwen_v2_moe  # noqa: F401  (registration)
    return POLICIES.get(model_type)
