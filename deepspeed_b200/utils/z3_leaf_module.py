"""ZeRO-3 leaf-module marking (reference: ``utils/z3_leaf_module.py``).

A "leaf" module is gathered/released as one unit even if it has children (needed for MoE blocks
whose experts execute in data-dependent order).
"""
from typing import List, Type

import torch

_FLAG = "_z3_leaf"


def z3_leaf_module(module: torch.nn.Module) -> bool:
    return getattr(module, _FLAG, False)


def z3_leaf_parameter(p) -> bool:
    return getattr(p, "ds_z3_leaf_module", None) is not None


def get_z3_leaf_modules(model: torch.nn.Module) -> List[torch.nn.Module]:
    return [m for m in model.modules() if z3_leaf_module(m)]


def _mark(model, classes, flag):
    assert all(isinstance(c, (type, str)) for c in classes), "leaf_module_classes must be classes or class names"
    hits = []
    for m in model.modules():
        for c in classes:
            if (isinstance(c, type) and isinstance(m, c)) or (isinstance(c, str) and type(m).__name__ == c):
                setattr(m, _FLAG, flag)
                hits.append(m)
                break
    if not hits:
        raise ValueError(f"no modules matching {classes} found in model")
    return hits


def set_z3_leaf_modules(model, leaf_module_classes: List[Type]):
    return _mark(model, leaf_module_classes, True)


def unset_z3_leaf_modules(model, leaf_module_classes: List[Type]):
    return _mark(model, leaf_module_classes, False)


def set_z3_leaf_module(model: torch.nn.Module, flag: bool):
    """Mark / unmark one module as a ZeRO-3 leaf (its children are gathered together with it)."""
    model._z3_leaf = bool(flag)
    for p in model.parameters():
        p._z3_leaf_param = bool(flag)
