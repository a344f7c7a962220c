"""Declarative parameters for layer containers (reference ``model_implementations/parameter_base.py``).

A ``ParameterBase`` subclass lists its *dependencies* as annotations (``params: torch.Tensor`` / ``ParamList``); the
owning container feeds checkpoint tensors into them and, once every dependency is set, ``finalize()`` produces the
tensor the model consumes (fused / transposed / sharded / quantised through ``inference_model.transform_*``).
"""
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


def ParametrizedList(count_attr: str) -> ParamList:
    """Reference spelling of a list dependency: ``experts = ParametrizedList("n_experts")``."""
    return ParamList(count_attr)


def make_param_getter(clsname, param):
    """Property getter of tensor dependency ``param``."""

    def getter(self):
        return self._deps[param]

    getter.__qualname__ = f"{clsname}.{param}"
    return getter


def make_param_setter(clsname, param):
    """Property setter: stores the dependency and finalises the parameter once every dependency is present."""

    def setter(self, value):
        self._deps[param] = value
        self._maybe_complete()

    return setter


def make_readonly_setter():
    """Setter installed on list dependencies: the list object is fixed, only its items are assigned."""

    def setter(self, value):
        raise ValueError("Cannot set a ParamList directly; assign its items (``param.experts[i] = tensor``)")

    return setter


def _make_list_getter(name):

    def getter(self):
        return _ListProxy(self, self._lists[name])

    return getter


class ParameterMetaclass(type):
    """Analyses a parameter class ONCE, at class creation: ``torch.Tensor`` annotations become tensor dependencies,
    ``ParamList`` attributes become list dependencies, and each gets a property (reference ``parameter_base.py:58``)."""

    def __new__(mcs, clsname, bases, attrs):
        tensors = [n for n, hint in attrs.get("__annotations__", {}).items() if hint is torch.Tensor or hint == "torch.Tensor"]
        lists = {n: v for n, v in attrs.items() if isinstance(v, ParamList)}
        for n in tensors:
            attrs[n] = property(make_param_getter(clsname, n), make_param_setter(clsname, n))
        for n in lists:
            attrs[n] = property(_make_list_getter(n), make_readonly_setter())
        cls = super().__new__(mcs, clsname, bases, attrs)
        inherited_t = [n for b in bases for n in getattr(b, "tensor_dependencies", ())]
        inherited_l = {n: v for b in bases for n, v in getattr(b, "list_dependencies", {}).items()}
        cls.tensor_dependencies = tuple(dict.fromkeys(inherited_t + tensors))
        cls.list_dependencies = {**inherited_l, **lists}
        cls.n_dependencies = len(cls.tensor_dependencies) + len(cls.list_dependencies)
        return cls


class ParameterBase(metaclass=ParameterMetaclass):

    def __init__(self, model=None, on_complete=None):
        self.inference_model = model
        self._on_complete = on_complete
        self._deps = {n: None for n in type(self).tensor_dependencies}
        self._lists = {n: _ListState(getattr(model, v.count_attr)) for n, v in type(self).list_dependencies.items()}
        self.result = None

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
