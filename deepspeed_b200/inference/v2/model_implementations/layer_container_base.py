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


class LayerContainer:
    PARAM_MAPPING = {}

    def __init__(self, model=None) -> None:
        self.inference_model = model
        self._params = {}
        self._finalized = set()
        for name, hint in get_type_hints(type(self)).items():
            if isinstance(hint, type) and issubclass(hint, ParameterBase):
                self._params[name] = hint(model, on_complete=lambda p, n=name: self._finalized.add(n))
        self._rules = []
        for src, targets in self.PARAM_MAPPING.items():
            rx = re.compile("^" + re.escape(src).replace("\\*", r"(\d+)") + "$")
            self._rules.append((rx, targets if isinstance(targets, (list, tuple)) else [targets]))

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
