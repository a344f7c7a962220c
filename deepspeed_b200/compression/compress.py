"""Entry points (reference ``compression/compress.py``: ``init_compression :100``, ``redundancy_clean :148``,
``student_initialization :192``)."""
import json
import re

import torch

from . import constants as C
from .config import get_compression_config
from .helper import compression_preparation, fix_compression, recursive_getattr


def check_deepspeed_config(config):
    if isinstance(config, dict):
        return config
    if isinstance(config, str):
        with open(config) as f:
            return json.load(f)
    raise ValueError(f"Expected a string path to an existing deepspeed config, or a dictionary. Received: {config}")


def get_module_name(group_name, model, key_word, exist_module_name, mpu=None, verbose=True):
    """Names of compressible modules whose qualified name matches the regex / substring ``key_word``."""
    from .helper import is_module_compressible
    found = []
    for name, module in model.named_modules():
        if name and is_module_compressible(module, mpu) and (key_word == "*" or re.search(key_word, name)):
            if name in exist_module_name and verbose:
                raise ValueError(f"{name} is already added to compression, please check your config file for {group_name}.")
            if name not in exist_module_name:
                exist_module_name = exist_module_name + [name] if isinstance(exist_module_name, list) else exist_module_name | {name}
                found.append(name)
    return found, exist_module_name


def get_compress_methods(model, compress_methods, mpu=None):
    """-> list of ``[module_name_lists, related_module_name_lists, {technique: params}]`` per config group."""
    out = []
    for method, mc in compress_methods.items():
        if method == C.LAYER_REDUCTION:
            continue
        shared = mc[C.SHARED_PARAMETERS]
        if not shared[C.TECHNIQUE_ENABLED]:
            continue
        seen = []
        for gname, g in mc[C.DIFFERENT_GROUPS].items():
            names, related = [], []
            for i, kw in enumerate(g[C.DIFFERENT_GROUPS_MODULE_SCOPE]):
                found, seen = get_module_name(gname, model, kw, seen, mpu=mpu)
                names.append(found)
                rel = g[C.DIFFERENT_GROUPS_RELATED_MODULE_SCOPE]
                if rel:
                    r_found = []
                    for rkw in rel[i] if isinstance(rel[i], (list, tuple)) else [rel[i]]:
                        rf, _ = get_module_name(gname, model, rkw, [], mpu=mpu, verbose=False)
                        r_found.append(rf)
                    related.append(r_found)
            params = dict(shared)
            params.update(g[C.DIFFERENT_GROUPS_PARAMETERS])
            out.append([names, related, {method: params}])
    return out


def init_compression(model, deepspeed_config, teacher_model=None, mpu=None):
    cfg = get_compression_config(check_deepspeed_config(deepspeed_config))
    c_model = model.module if hasattr(model, "module") else model
    if cfg[C.LAYER_REDUCTION][C.LAYER_REDUCTION_ENABLED]:
        assert teacher_model is not None, "Teacher model is required for layer reduction"
        student_initialization(c_model, teacher_model, deepspeed_config)
    compression_preparation(c_model, get_compress_methods(c_model, cfg, mpu=mpu), mpu)
    sp = cfg[C.SPARSE_PRUNING][C.SHARED_PARAMETERS]
    if sp[C.TECHNIQUE_ENABLED] and sp[C.SPARSE_PRUNING_METHOD] == C.SPARSE_PRUNING_METHOD_SNIP_MOMENTUM:
        # block-sparse SNIP-momentum pruning: masks are driven by hooks on the model (step begin) and on the optimizer
        # (``rewrite_optimizer_step(optimizer).pruners = model.pruners``; the engine does this when it builds the optimizer)
        from .helper import generate_pruners, register_on_step_begin
        c_model.pruners = generate_pruners(
            {"target_sparsity": 1 - sp.get(C.SPARSE_PRUNING_DENSE_RATIO, 0.1), "pattern": sp.get(C.SPARSE_PRUNING_BLOCK_PATTERN, "4x1"),
             "pruning_frequency": sp.get(C.SPARSE_PRUNING_SCHEDULE_OFFSET_STRIDE, 1), "start_step": sp[C.TECHNIQUE_SCHEDULE_OFFSET],
             "end_step": sp.get(C.TECHNIQUE_SCHEDULE_OFFSET_END, sp[C.TECHNIQUE_SCHEDULE_OFFSET]),
             "excluded_op_names": sp.get(C.SPARSE_PRUNING_EXCLUDED_MODULES, [])}, c_model)
        c_model._pruner_hook = register_on_step_begin(c_model)
    return model


def redundancy_clean(model, deepspeed_config, mpu=None):
    """Make compression permanent (quantised values written, pruned rows/heads/channels physically removed)."""
    cfg = get_compression_config(check_deepspeed_config(deepspeed_config))
    c_model = model.module if hasattr(model, "module") else model
    order = [C.WEIGHT_QUANTIZATION, C.SPARSE_PRUNING, C.ROW_PRUNING, C.HEAD_PRUNING, C.CHANNEL_PRUNING,
             C.ACTIVATION_QUANTIZATION]
    layers = sorted(get_compress_methods(c_model, cfg, mpu=mpu), key=lambda x: order.index(next(iter(x[2]))))
    for names, related, technique in layers:
        stored = []
        need_mask = bool(related)
        for i, group in enumerate(names):
            for j, name in enumerate(group):
                mask = fix_compression(c_model, name, technique, dim_reduction=need_mask)
                if need_mask:
                    stored.append((i, mask))
        if need_mask:
            k = 0
            for i, group in enumerate(names):
                for j, name in enumerate(group):
                    mask = stored[k][1]
                    k += 1
                    for rlist in related[i]:
                        if j < len(rlist):
                            fix_compression(c_model, rlist[j], technique, mask=mask, dim_reduction=True)
    return model


def student_initialization(student_model, teacher_model, deepspeed_config):
    """Layer reduction: copy ``teacher_layer[i]`` into student layer ``i`` and the listed other modules."""
    cfg = get_compression_config(check_deepspeed_config(deepspeed_config))[C.LAYER_REDUCTION]
    prefix = cfg[C.MODULE_NAME_PREFIX]
    teacher_layers = cfg[C.TEACHER_LAYER]
    assert len(teacher_layers) == cfg[C.KEEP_NUMBER_LAYER]
    with torch.no_grad():
        for s_idx, t_idx in enumerate(teacher_layers):
            s = recursive_getattr(student_model, f"{prefix}.{s_idx}")
            t = recursive_getattr(teacher_model, f"{prefix}.{t_idx}")
            for sp, tp in zip(s.parameters(), t.parameters()):
                sp.data.copy_(tp.data)
        for name in cfg.get(C.OTHER_MODULE_NAME, []):
            s, t = recursive_getattr(student_model, name), recursive_getattr(teacher_model, name)
            for sp, tp in zip(s.parameters(), t.parameters()):
                sp.data.copy_(tp.data)
