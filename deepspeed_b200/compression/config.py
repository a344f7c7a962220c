"""Normalise the ``compression_training`` block: every technique gets ``shared_parameters`` (with defaults) and
``different_groups`` (each with ``params`` / ``modules`` / ``related_modules``).  Reference ``compression/config.py``."""
import copy

from . import constants as C

_SHARED_DEFAULTS = {
    C.WEIGHT_QUANTIZATION: {C.TECHNIQUE_ENABLED: False, C.WEIGHT_QUANTIZE_KERNEL: False, C.TECHNIQUE_SCHEDULE_OFFSET: 0,
                            C.WEIGHT_QUANTIZE_GROUPS: 1, C.WEIGHT_QUANTIZE_VERBOSE: False,
                            C.WEIGHT_QUANTIZE_TYPE: C.WEIGHT_QUANTIZE_SYMMETRIC,
                            C.WEIGHT_QUANTIZE_IN_FORWARD_ENABLED: False,
                            C.WEIGHT_QUANTIZE_ROUNDING: C.WEIGHT_QUANTIZE_NEAREST_ROUNDING,
                            C.WEIGHT_QUANTIZE_FP16_MIXED_QUANTIZE: {C.TECHNIQUE_ENABLED: False,
                                                                    C.WEIGHT_QUANTIZE_CHANGE_RATIO: 0.001}},
    C.ACTIVATION_QUANTIZATION: {C.TECHNIQUE_ENABLED: False, C.ACTIVATION_QUANTIZE_TYPE: "symmetric",
                                C.ACTIVATION_QUANTIZE_RANGE: C.ACTIVATION_QUANTIZE_RANGE_DYNAMIC,
                                C.TECHNIQUE_SCHEDULE_OFFSET: 1000},
    C.SPARSE_PRUNING: {C.TECHNIQUE_ENABLED: False, C.SPARSE_PRUNING_METHOD: C.SPARSE_PRUNING_METHOD_L1,
                       C.TECHNIQUE_SCHEDULE_OFFSET: 1000},
    C.ROW_PRUNING: {C.TECHNIQUE_ENABLED: False, C.ROW_PRUNING_METHOD: "l1", C.TECHNIQUE_SCHEDULE_OFFSET: 1000},
    C.HEAD_PRUNING: {C.TECHNIQUE_ENABLED: False, C.HEAD_PRUNING_METHOD: "topk", C.TECHNIQUE_SCHEDULE_OFFSET: 1000},
    C.CHANNEL_PRUNING: {C.TECHNIQUE_ENABLED: False, C.CHANNEL_PRUNING_METHOD: "l1", C.TECHNIQUE_SCHEDULE_OFFSET: 1000},
}


def _technique(block, name):
    sub = copy.deepcopy(block.get(name, {}))
    shared = copy.deepcopy(_SHARED_DEFAULTS[name])
    shared.update(sub.get(C.SHARED_PARAMETERS, {}))
    if name == C.HEAD_PRUNING and shared[C.TECHNIQUE_ENABLED]:
        assert C.HEAD_PRUNING_NUM_HEADS in shared, "head_pruning requires shared_parameters.num_heads"
    groups = {}
    for gname, g in sub.get(C.DIFFERENT_GROUPS, {}).items():
        assert C.DIFFERENT_GROUPS_PARAMETERS in g, f"group {gname} of {name} needs 'params'"
        groups[gname] = {
            C.DIFFERENT_GROUPS_PARAMETERS: dict(g[C.DIFFERENT_GROUPS_PARAMETERS]),
            C.DIFFERENT_GROUPS_MODULE_SCOPE: g.get(C.DIFFERENT_GROUPS_MODULE_SCOPE, [C.DIFFERENT_GROUPS_MODULE_SCOPE_DEFAULT]),
            C.DIFFERENT_GROUPS_RELATED_MODULE_SCOPE: g.get(C.DIFFERENT_GROUPS_RELATED_MODULE_SCOPE,
                                                           C.DIFFERENT_GROUPS_RELATED_MODULE_SCOPE_DEFAULT),
        }
        p = groups[gname][C.DIFFERENT_GROUPS_PARAMETERS]
        if name == C.WEIGHT_QUANTIZATION:
            assert C.WEIGHT_QUANTIZE_START_BITS in p and C.WEIGHT_QUANTIZE_TARGET_BITS in p
            p.setdefault(C.WEIGHT_QUANTIZATION_PERIOD, 1)
        elif name == C.ACTIVATION_QUANTIZATION:
            assert C.ACTIVATION_QUANTIZE_BITS in p
        elif not (name == C.SPARSE_PRUNING and shared.get(C.SPARSE_PRUNING_METHOD) == C.SPARSE_PRUNING_METHOD_SNIP_MOMENTUM):
            assert "dense_ratio" in p, f"{name} group {gname} requires dense_ratio"
    return {C.SHARED_PARAMETERS: shared, C.DIFFERENT_GROUPS: groups}


def get_compression_config(param_dict):
    block = param_dict.get(C.COMPRESSION_TRAINING, {})
    out = {t: _technique(block, t) for t in C.TECHNIQUES}
    lr = copy.deepcopy(block.get(C.LAYER_REDUCTION, {}))
    lr.setdefault(C.LAYER_REDUCTION_ENABLED, False)
    out[C.LAYER_REDUCTION] = lr
    return out


# ---- per-technique accessors (reference ``compression/config.py:get_*``) -------------------------------------------------
def _shared(key, block):
    return _technique({key: block}, key)[C.SHARED_PARAMETERS]


def _groups(key, block):
    return _technique({key: block}, key)[C.DIFFERENT_GROUPS]


# ``get_<t>(compression_block)`` -> normalised technique; the two others take the technique's own block.
def get_weight_quantization(param_dict):
    return _technique(param_dict, C.WEIGHT_QUANTIZATION)


def get_weight_quantization_shared_parameters(param_dict):
    return _shared(C.WEIGHT_QUANTIZATION, param_dict)


def get_weight_quantization_different_groups(param_dict):
    return _groups(C.WEIGHT_QUANTIZATION, param_dict)


def get_activation_quantization(param_dict):
    return _technique(param_dict, C.ACTIVATION_QUANTIZATION)


def get_activation_quantization_shared_parameters(param_dict):
    return _shared(C.ACTIVATION_QUANTIZATION, param_dict)


def get_activation_quantization_different_groups(param_dict):
    return _groups(C.ACTIVATION_QUANTIZATION, param_dict)


def get_sparse_pruning(param_dict):
    return _technique(param_dict, C.SPARSE_PRUNING)


def get_sparse_pruning_shared_parameters(param_dict):
    return _shared(C.SPARSE_PRUNING, param_dict)


def get_sparse_pruning_different_groups(param_dict):
    return _groups(C.SPARSE_PRUNING, param_dict)


def get_row_pruning(param_dict):
    return _technique(param_dict, C.ROW_PRUNING)


def get_row_pruning_shared_parameters(param_dict):
    return _shared(C.ROW_PRUNING, param_dict)


def get_row_pruning_different_groups(param_dict):
    return _groups(C.ROW_PRUNING, param_dict)


def get_head_pruning(param_dict):
    return _technique(param_dict, C.HEAD_PRUNING)


def get_head_pruning_shared_parameters(param_dict):
    return _shared(C.HEAD_PRUNING, param_dict)


def get_head_pruning_different_groups(param_dict):
    return _groups(C.HEAD_PRUNING, param_dict)


def get_channel_pruning(param_dict):
    return _technique(param_dict, C.CHANNEL_PRUNING)


def get_channel_pruning_shared_parameters(param_dict):
    return _shared(C.CHANNEL_PRUNING, param_dict)


def get_channel_pruning_different_groups(param_dict):
    return _groups(C.CHANNEL_PRUNING, param_dict)


def get_layer_reduction(param_dict):
    """``param_dict``: the ``compression_training`` block -> the layer-reduction section with ``enabled`` defaulted."""
    lr = copy.deepcopy(param_dict.get(C.LAYER_REDUCTION, {}))
    lr.setdefault(C.LAYER_REDUCTION_ENABLED, False)
    return lr


def get_layer_reduction_enabled(param_dict):
    return bool(param_dict.get(C.LAYER_REDUCTION, {}).get(C.LAYER_REDUCTION_ENABLED, False))


def get_layer_reduction_params(param_dict):
    lr = copy.deepcopy(param_dict.get(C.LAYER_REDUCTION, {}))
    lr.pop(C.LAYER_REDUCTION_ENABLED, None)
    return lr or False


def get_quantize_enabled(param_dict):
    """Is weight quantisation switched on in the ``compression_training`` block?"""
    return bool(param_dict.get(C.COMPRESSION_TRAINING, param_dict).get(C.WEIGHT_QUANTIZATION, {}).get(C.SHARED_PARAMETERS, {}).get(
        C.TECHNIQUE_ENABLED, False))
