"""Declarative parameters for layer containers (reference ``model_implementations/parameter_base.py``).

A ``ParameterBase`` subclass lists its *dependencies* as annotations (``params: torch.Tensor`` / ``ParamList``); the
owning container feeds checkpoint tensors into them and, once every dependency is set, ``finalize()`` produces the
tensor the model consumes (fused / transposed / sharded / quantised through ``inference_model.transform_*``).
"""
from typing import get_type_hints

import torch


class ParamList:
    """A fixed-length list dependency (e.g. one tensor per expert); length = ``getattr(inference_model, count_attr)``."""

    def __init__(self, count_attr: str):
        self.count_attr = count_attr


class _ListState:

    def __init__(self, n):
        self.items = [None] * n

    def __setitem__(self, i, v):
        self.items[i] = v

    def __getitem__(self, i):
        return self.items[i]

    def __len__(self):
        return len(self.items)

    @property
    def complete(self):
        return all(x is not None for x in self.items)


class ParameterBase:

    def __init__(self, model=None, on_complete=None):
        self.inference_model = model
        self._on_complete = on_complete
        self._deps = {}
        self._lists = {}
        for name, hint in get_type_hints(type(self)).items():
            if hint is torch.Tensor:
                self._deps[name] = None
        for name, val in vars(type(self)).items():
            if isinstance(val, ParamList):
                n = getattr(model, val.count_attr)
                self._lists[name] = _ListState(n)
        self.result = None

    def __setattr__(self, key, value):
        if key not in ("_deps", "_lists") and "_deps" in self.__dict__ and key in self._deps:
            self._deps[key] = value
            self._maybe_complete()
            return
        super().__setattr__(key, value)

    def __getattribute__(self, key):
        d = object.__getattribute__(self, "__dict__")
        if "_deps" in d and key in d["_deps"]:
            return d["_deps"][key]
        if "_lists" in d and key in d["_lists"]:
            return _ListProxy(self, d["_lists"][key])
        return object.__getattribute__(self, key)

    @property
    def complete(self):
        return all(v is not None for v in self._deps.values()) and all(s.complete for s in self._lists.values())

    def _maybe_complete(self):
        if self.complete and self.result is None:
            self.result = self.finalize()
            if self._on_complete is not None:
                self._on_complete(self)

    def finalize(self) -> torch.Tensor:
        raise NotImplementedError


class _ListProxy:

    def __init__(self, owner, state):
        self._owner, self._state = owner, state

    def __setitem__(self, i, v):
        self._state[i] = v
        self._owner._maybe_complete()

    def __getitem__(self, i):
        return self._state[i]

    def __len__(self):
        return len(self._state)

    def __iter__(self):
        return iter(self._state.items)
