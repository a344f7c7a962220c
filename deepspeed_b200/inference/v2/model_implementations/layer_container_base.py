"""Layer containers: named groups of ``ParameterBase`` objects filled from checkpoint names (reference
``model_implementations/layer_container_base.py``).

A subclass declares its parameters as annotations and a ``PARAM_MAPPING`` from checkpoint suffixes to
``"<param>.<dependency>"`` targets (a list target feeds several parameters; ``"name.*"`` wildcards feed list
dependencies by index).  ``set_dependency(name, tensor)`` routes one checkpoint tensor; ``is_initialized`` turns true when
every parameter finalised.
"""
import re
from typing import get_type_hints

import torch

from .parameter_base import ParameterBase


def make_finalization_callback(all_names):
    """Build the callback a container hands to its parameters: marks the parameter (looked up by identity under any of
    ``all_names``) as finalised."""

    def finalization_callback(self, param: ParameterBase, finalized_param=None) -> None:
        for name in all_names:
            if self._params.get(name) is param:
                self._finalized.add(name)

    return finalization_callback


class LayerMetaclass(type):
    """Analyses a container class once: which annotations are ``ParameterBase`` types (``annotation_attrs``), the compiled
    ``PARAM_MAPPING`` routing rules, and the finalisation callback (reference ``layer_container_base.py:42``)."""

    def __new__(mcs, clsname, bases, attrs):
        cls = super().__new__(mcs, clsname, bases, attrs)
        try:
            hints = get_type_hints(cls)
        except Exception:
            hints = dict(attrs.get("__annotations__", {}))
        cls.annotation_attrs = {n: h for n, h in hints.items() if isinstance(h, type) and issubclass(h, ParameterBase)}
        rules = []
        for src, targets in getattr(cls, "PARAM_MAPPING", {}).items():
            rx = re.compile("^" + re.escape(src).replace("\\*", r"(\d+)") + "$")
            rules.append((rx, list(targets) if isinstance(targets, (list, tuple)) else [targets]))
        cls._compiled_rules = rules
        cls._finalization_callback = make_finalization_callback(tuple(cls.annotation_attrs))
        return cls


class LayerContainer(metaclass=LayerMetaclass):
    PARAM_MAPPING = {}

    def __init__(self, model=None) -> None:
        self.inference_model = model
        self._params = {}
        self._finalized = set()
        for name, ptype in type(self).annotation_attrs.items():
            self._params[name] = ptype(model, on_complete=lambda p: type(self)._finalization_callback(self, p))
        self._rules = type(self)._compiled_rules

    def __getattr__(self, name):
        params = self.__dict__.get("_params", {})
        if name in params:
            p = params[name]
            return p.result if p.result is not None else p
        raise AttributeError(name)

    @property
    def n_params(self) -> int:
        return len(self._params)

    @property
    def is_initialized(self) -> bool:
        return len(self._finalized) == len(self._params)

    @property
    def is_populated(self) -> bool:
        return self.is_initialized

    def set_dependency(self, dep_name: str, dep_value: torch.Tensor) -> bool:
        """Route a checkpoint tensor (suffix relative to the layer) to its parameter dependency; False if unmapped."""
        for rx, targets in self._rules:
            m = rx.match(dep_name)
            if m is None:
                continue
            for t in targets:
                pname, dname = t.split(".", 1)
                param = self._params[pname]
                if m.groups():
                    getattr(param, dname)[int(m.group(1))] = dep_value
                else:
                    setattr(param, dname, dep_value)
            return True
        return False

    def direct_injection(self, name: str, tensor: torch.Tensor) -> None:
        """Bypass dependencies: ``tensor`` is already in its final form (flattened-model restore)."""
        self._params[name].result = tensor
        self._finalized.add(name)

    def parameters(self):
        return {n: p.result for n, p in self._params.items()}
