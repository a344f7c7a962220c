"""Typed-config base class and helpers.

Parity target: reference ``runtime/config_utils.py:16 DeepSpeedConfigModel`` -- pydantic model
that (a) strips ``"auto"`` values so defaults apply, (b) remaps deprecated fields onto their
replacements, (c) forbids unknown keys.  Deprecated fields are declared with
``json_schema_extra={"deprecated": True, "new_param": "a.b", "new_param_fn": fn}``.
"""
import collections
import json
from functools import reduce
from typing import Any

from pydantic import BaseModel, ConfigDict

from deepspeed_b200.utils.logging import logger


class DeprecatedParameterConflict(AssertionError, ValueError):
    """A deprecated key and its replacement were both given (the reference asserts, ``config_utils.py:89``)."""


class DeepSpeedConfigModel(BaseModel):
    model_config = ConfigDict(
        validate_default=True,
        validate_assignment=True,
        use_enum_values=True,
        populate_by_name=True,
        extra="forbid",
        arbitrary_types_allowed=True,
        protected_namespaces=(),
    )

    def __init__(self, strict=False, **data):
        if not strict:  # "auto" == use the default
            data = {k: v for k, v in data.items() if not (isinstance(v, str) and v == "auto")}
        super().__init__(**data)
        self._apply_deprecations()

    def _apply_deprecations(self):
        cls = type(self)
        for name, field in cls.model_fields.items():
            extra = field.json_schema_extra or {}
            if not isinstance(extra, dict) or not extra.get("deprecated", False):
                continue
            if name not in self.model_fields_set:
                continue
            new_param = extra.get("new_param", "")
            msg = f"Config parameter {name} is deprecated" + (f", use {new_param} instead" if new_param else "")
            dep_msg = extra.get("deprecated_msg", "")
            logger.warning(msg + (f". {dep_msg}" if dep_msg else ""))
            if not new_param or not extra.get("set_new_param", True):
                continue
            fn = extra.get("new_param_fn", lambda x: x)
            value = fn(getattr(self, name))
            path = new_param.split(".")
            target = reduce(getattr, path[:-1], self)
            leaf = path[-1]
            if leaf in getattr(target, "model_fields_set", ()):  # both old and new given
                raise DeprecatedParameterConflict(f"Cannot provide deprecated parameter '{name}' and replacing parameter "
                                 f"'{new_param}' together")
            try:
                setattr(target, leaf, value)
            except Exception as e:  # pragma: no cover
                logger.error(f"Tried setting value for '{new_param}' with value from deprecated '{name}'")
                raise e

    def __repr__(self):
        return f"{type(self).__name__}({self.model_dump()})"


def get_config_default(config, field_name):
    assert field_name in config.model_fields, f"'{field_name}' is not a field in {config}"
    assert not config.model_fields.get(field_name).is_required(), f"'{field_name}' is a required field"
    return config.model_fields.get(field_name).get_default()


class pp_int(int):
    """int that prints with thousands separators (used for bucket-size defaults in dumps)."""

    def __new__(cls, val, custom_print_str=None):
        inst = super().__new__(cls, val)
        inst.custom_print_str = custom_print_str
        return inst

    def __repr__(self):
        if self.custom_print_str:
            return self.custom_print_str
        return f"{self.real:,}"


class ScientificNotationEncoder(json.JSONEncoder):
    """json encoder printing large numbers as 1e+09 (config dump readability)."""

    def iterencode(self, o, _one_shot=False, level=0):
        indent = self.indent if self.indent is not None else 4
        prefix_close = " " * level * indent
        level += 1
        prefix = " " * level * indent
        if isinstance(o, bool):
            return "true" if o else "false"
        elif isinstance(o, (float, int)):
            if o > 1e3:
                return f"{o:e}"
            return f"{o}"
        elif isinstance(o, collections.abc.Mapping):
            x = [f'\n{prefix}"{k}": {self.iterencode(v, level=level)}' for k, v in o.items()]
            return "{" + ", ".join(x) + f"\n{prefix_close}" + "}"
        elif isinstance(o, collections.abc.Sequence) and not isinstance(o, str):
            return f"[{', '.join(map(self.iterencode, o))}]"
        return "\n, ".join(super().iterencode(o, _one_shot))


def get_scalar_param(param_dict, name, default):
    return param_dict.get(name, default)


def get_list_param(param_dict, name, default):
    return param_dict.get(name, default)


def get_dict_param(param_dict, name, default):
    return param_dict.get(name, default)


def dict_raise_error_on_duplicate_keys(ordered_pairs):
    """``object_pairs_hook`` for json/hjson loads that rejects duplicate keys."""
    d = dict(ordered_pairs)
    if len(d) != len(ordered_pairs):
        counts = collections.Counter(k for k, _ in ordered_pairs)
        dups = [k for k, c in counts.items() if c > 1]
        raise ValueError(f"Duplicate keys in DeepSpeed config: {dups}")
    return d


def as_plain(obj: Any):
    if isinstance(obj, BaseModel):
        return obj.model_dump()
    return obj


class DeepSpeedConfigObject:
    """Legacy (non-pydantic) config base: JSON repr of the instance dict (reference ``config_utils.py``)."""

    def repr(self):
        return self.__dict__

    def __repr__(self):
        return json.dumps(self.__dict__, sort_keys=True, indent=4, cls=ScientificNotationEncoder)
