"""MoE parameter helpers (reference ``moe/utils.py:72`` param-group split)."""
from typing import Dict, List

import torch


def has_moe_layers(m):
    from .layer import MoE
    n = [mod.num_experts for mod in m.modules() if isinstance(mod, MoE)]
    return len(n) > 0, n


def is_moe_param(param: torch.Tensor) -> bool:
    return hasattr(param, "allreduce") and not param.allreduce


def split_params_into_shared_and_expert_params(params):
    shared, expert = [], []
    for p in params:
        (expert if is_moe_param(p) else shared).append(p)
    return shared, expert


def split_params_grads_into_shared_and_expert_params(group):
    shared, expert = [], []
    for p in group:
        if p.grad is not None:
            (expert if is_moe_param(p) else shared).append(p.grad.to(p.dtype))
    return shared, expert


def split_params_into_different_moe_groups_for_optimizer(param_groups, max_group_size=178956971):
    """Move expert parameters of every group into per-``group_name`` groups flagged ``moe=True``."""
    if isinstance(param_groups, tuple):
        param_groups = list(param_groups)
    elif isinstance(param_groups, dict):
        param_groups = [param_groups]
    elif not isinstance(param_groups, list):
        raise ValueError(f"Unknown param group type of {type(param_groups)}")
    names = {p.group_name for g in param_groups for p in g["params"] if is_moe_param(p)}
    out = []
    moe_groups: Dict[str, Dict] = {}
    for g in param_groups:
        new = {k: v for k, v in g.items() if k != "params"}
        new["params"] = [p for p in g["params"] if not is_moe_param(p)]
        out.append(new)
        for name in sorted(names):
            ps = [p for p in g["params"] if is_moe_param(p) and p.group_name == name]
            if not ps:
                continue
            mg = {k: v for k, v in g.items() if k != "params"}
            mg.update({"name": name, "moe": True, "params": ps})
            out.append(mg)
    return out


def is_moe_param_group(param_group):
    return param_group.get("moe", False)


def configure_moe_param_groups(model_parameters: List):
    assert isinstance(model_parameters, list)
    if model_parameters and isinstance(model_parameters[0], dict):
        return split_params_into_different_moe_groups_for_optimizer(model_parameters)
    return split_params_into_different_moe_groups_for_optimizer([{"params": list(model_parameters)}])
