"""Config keys of ``compression_training`` (reference ``compression/constants.py``).

Names are declared section by section through ``_declare`` (later sections may refer to earlier names)."""


def _declare(**names):
    globals().update(names)
    return names


_declare(
    COMPRESSION_TRAINING="compression_training",
    SHARED_PARAMETERS="shared_parameters",
    DIFFERENT_GROUPS="different_groups",
    TECHNIQUE_ENABLED="enabled",
    TECHNIQUE_SCHEDULE_OFFSET="schedule_offset",
    TECHNIQUE_SCHEDULE_OFFSET_END="schedule_offset_end",
)

_declare(
    DIFFERENT_GROUPS_PARAMETERS="params",
    DIFFERENT_GROUPS_MODULE_SCOPE="modules",
    DIFFERENT_GROUPS_MODULE_SCOPE_DEFAULT="*",
    DIFFERENT_GROUPS_RELATED_MODULE_SCOPE="related_modules",
    DIFFERENT_GROUPS_RELATED_MODULE_SCOPE_DEFAULT=None,
)

_declare(
    LAYER_REDUCTION="layer_reduction",
    LAYER_REDUCTION_ENABLED="enabled",
    KEEP_NUMBER_LAYER="keep_number_layer",
    MODULE_NAME_PREFIX="module_name_prefix",
)

_declare(
    TEACHER_LAYER="teacher_layer",
    OTHER_MODULE_NAME="other_module_name",
    WEIGHT_QUANTIZATION="weight_quantization",
    WEIGHT_QUANTIZATION_PERIOD="quantization_period",
    WEIGHT_QUANTIZE_IN_FORWARD_ENABLED="quantize_weight_in_forward",
    WEIGHT_QUANTIZE_KERNEL="quantizer_kernel",
    WEIGHT_QUANTIZE_GROUPS="quantize_groups",
    WEIGHT_QUANTIZE_VERBOSE="quantize_verbose",
    WEIGHT_QUANTIZE_TYPE="quantization_type",
    WEIGHT_QUANTIZE_SYMMETRIC="symmetric",
    WEIGHT_QUANTIZE_ASYMMETRIC="asymmetric",
    WEIGHT_QUANTIZE_ROUNDING="rounding",
    WEIGHT_QUANTIZE_STOCHASTIC_ROUNDING="stochastic",
    WEIGHT_QUANTIZE_NEAREST_ROUNDING="nearest",
    WEIGHT_QUANTIZE_FP16_MIXED_QUANTIZE="fp16_mixed_quantize",
    WEIGHT_QUANTIZE_CHANGE_RATIO="quantize_change_ratio",
    WEIGHT_QUANTIZE_START_BITS="start_bits",
    WEIGHT_QUANTIZE_TARGET_BITS="target_bits",
)

_declare(
    ACTIVATION_QUANTIZATION="activation_quantization",
    ACTIVATION_QUANTIZE_TYPE="quantization_type",
    ACTIVATION_QUANTIZE_RANGE="range_calibration",
    ACTIVATION_QUANTIZE_RANGE_STATIC="static",
    ACTIVATION_QUANTIZE_RANGE_DYNAMIC="dynamic",
    ACTIVATION_QUANTIZE_BITS="bits",
)

_declare(
    SPARSE_PRUNING="sparse_pruning",
    SPARSE_PRUNING_METHOD="method",
    SPARSE_PRUNING_METHOD_L1="l1",
    SPARSE_PRUNING_METHOD_TOPK="topk",
    SPARSE_PRUNING_METHOD_SNIP_MOMENTUM="snip_momentum",
    SPARSE_PRUNING_DENSE_RATIO="dense_ratio",
    SPARSE_PRUNING_BLOCK_PATTERN="block_pattern",
    SPARSE_PRUNING_SCHEDULE_OFFSET_STRIDE="schedule_offset_stride",
    SPARSE_PRUNING_EXCLUDED_MODULES="excluded_modules",
)

_declare(
    ROW_PRUNING="row_pruning",
    ROW_PRUNING_METHOD="method",
    ROW_PRUNING_DENSE_RATIO="dense_ratio",
    HEAD_PRUNING="head_pruning",
    HEAD_PRUNING_METHOD="method",
    HEAD_PRUNING_NUM_HEADS="num_heads",
    HEAD_PRUNING_DENSE_RATIO="dense_ratio",
)

_declare(
    CHANNEL_PRUNING="channel_pruning",
    CHANNEL_PRUNING_METHOD="method",
    CHANNEL_PRUNING_DENSE_RATIO="dense_ratio",
)

_declare(
    TECHNIQUES=(WEIGHT_QUANTIZATION, ACTIVATION_QUANTIZATION, SPARSE_PRUNING, ROW_PRUNING, HEAD_PRUNING, CHANNEL_PRUNING),
)
