"""Names of the top-level config keys and their defaults (role of reference ``runtime/constants.py``).

The authoritative schema is the pydantic models in ``runtime/config.py``; this module gives code written against the
reference's symbolic names (``from deepspeed.runtime.constants import TRAIN_BATCH_SIZE``) the same spellings.
"""
# ---- batch geometry ------------------------------------------------------------------------------------------------
TRAIN_BATCH_SIZE, TRAIN_BATCH_SIZE_DEFAULT = "train_batch_size", None
TRAIN_MICRO_BATCH_SIZE_PER_GPU, TRAIN_MICRO_BATCH_SIZE_PER_GPU_DEFAULT = "train_micro_batch_size_per_gpu", None
GRADIENT_ACCUMULATION_STEPS, GRADIENT_ACCUMULATION_STEPS_DEFAULT = "gradient_accumulation_steps", None
STEPS_PER_PRINT, STEPS_PER_PRINT_DEFAULT = "steps_per_print", None
DATALOADER_DROP_LAST, DATALOADER_DROP_LAST_DEFAULT = "dataloader_drop_last", False

# ---- optimizer / scheduler -------------------------------------------------------------------------------------------
OPTIMIZER, OPTIMIZER_TYPE_DEFAULT, OPTIMIZER_PARAMS = "optimizer", None, "params"
SCHEDULER, SCHEDULER_TYPE_DEFAULT, SCHEDULER_PARAMS = "scheduler", None, "params"
TYPE = "type"
LEGACY_FUSION, LEGACY_FUSION_DEFAULT = "legacy_fusion", False
MAX_GRAD_NORM = "max_grad_norm"
ZERO_ALLOW_UNTESTED_OPTIMIZER, ZERO_ALLOW_UNTESTED_OPTIMIZER_DEFAULT = "zero_allow_untested_optimizer", False
ZERO_FORCE_DS_CPU_OPTIMIZER, ZERO_FORCE_DS_CPU_OPTIMIZER_DEFAULT = "zero_force_ds_cpu_optimizer", True

# ---- precision -----------------------------------------------------------------------------------------------------------
BFLOAT16, BFLOAT16_OLD = "bf16", "bfloat16"
BFLOAT16_ENABLED, BFLOAT16_ENABLED_DEFAULT = "enabled", False
BFLOAT16_IMMEDIATE_GRAD_UPDATE, BFLOAT16_IMMEDIATE_GRAD_UPDATE_DEFAULT = "immediate_grad_update", False
FP16 = "fp16"
FP16_ENABLED, FP16_ENABLED_DEFAULT = "enabled", False
FP16_LOSS_SCALE, FP16_LOSS_SCALE_DEFAULT = "loss_scale", 0
FP16_AUTO_CAST, FP16_AUTO_CAST_DEFAULT = "auto_cast", False
FP16_INITIAL_SCALE_POWER, FP16_INITIAL_SCALE_POWER_DEFAULT = "initial_scale_power", 16
FP16_LOSS_SCALE_WINDOW, FP16_LOSS_SCALE_WINDOW_DEFAULT = "loss_scale_window", 1000
FP16_HYSTERESIS, FP16_HYSTERESIS_DEFAULT = "hysteresis", 2
FP16_CONSECUTIVE_HYSTERESIS, FP16_CONSECUTIVE_HYSTERESIS_DEFAULT = "consecutive_hysteresis", False
FP16_MIN_LOSS_SCALE, FP16_MIN_LOSS_SCALE_DEFAULT = "min_loss_scale", 1
FP16_MASTER_WEIGHTS_AND_GRADS, FP16_MASTER_WEIGHTS_AND_GRADS_DEFAULT = "fp16_master_weights_and_grads", False
AMP, AMP_ENABLED, AMP_ENABLED_DEFAULT = "amp", "enabled", False
DATA_TYPES, GRAD_ACCUM_DTYPE, GRAD_ACCUM_DTYPE_DEFAULT = "data_types", "grad_accum_dtype", None
COMMUNICATION_DATA_TYPE, COMMUNICATION_DATA_TYPE_DEFAULT = "communication_data_type", None
SEQ_PARALLEL_COMMUNICATION_DATA_TYPE, SEQ_PARALLEL_COMMUNICATION_DATA_TYPE_DEFAULT = "seq_parallel_communication_data_type", "fp32"

# ---- gradients ---------------------------------------------------------------------------------------------------------
GRADIENT_CLIPPING, GRADIENT_CLIPPING_DEFAULT = "gradient_clipping", 0.0
PRESCALE_GRADIENTS, PRESCALE_GRADIENTS_DEFAULT = "prescale_gradients", False
GRADIENT_PREDIVIDE_FACTOR, GRADIENT_PREDIVIDE_FACTOR_DEFAULT = "gradient_predivide_factor", 1.0
SPARSE_GRADIENTS, SPARSE_GRADIENTS_DEFAULT = "sparse_gradients", False
DISABLE_ALLGATHER, DISABLE_ALLGATHER_DEFAULT = "disable_allgather", False
GRAPH_HARVESTING, GRAPH_HARVESTING_DEFAULT = "graph_harvesting", False

# ---- diagnostics ---------------------------------------------------------------------------------------------------------
DUMP_STATE, DUMP_STATE_DEFAULT = "dump_state", False
VOCABULARY_SIZE, VOCABULARY_SIZE_DEFAULT = "vocabulary_size", None
WALL_CLOCK_BREAKDOWN, WALL_CLOCK_BREAKDOWN_DEFAULT = "wall_clock_breakdown", False
MEMORY_BREAKDOWN, MEMORY_BREAKDOWN_DEFAULT = "memory_breakdown", False

# ---- curvature / layer drop --------------------------------------------------------------------------------------------
EIGENVALUE = "eigenvalue"
EIGENVALUE_ENABLED, EIGENVALUE_ENABLED_DEFAULT = "enabled", False
EIGENVALUE_VERBOSE, EIGENVALUE_VERBOSE_DEFAULT = "verbose", False
EIGENVALUE_MAX_ITER, EIGENVALUE_MAX_ITER_DEFAULT = "max_iter", 100
EIGENVALUE_TOL, EIGENVALUE_TOL_DEFAULT = "tol", 1e-2
EIGENVALUE_STABILITY, EIGENVALUE_STABILITY_DEFAULT = "stability", 1e-6
EIGENVALUE_GAS_BOUNDARY_RESOLUTION, EIGENVALUE_GAS_BOUNDARY_RESOLUTION_DEFAULT = "gas_boundary_resolution", 1
EIGENVALUE_LAYER_NAME, EIGENVALUE_LAYER_NAME_DEFAULT = "layer_name", "bert.encoder.layer"
EIGENVALUE_LAYER_NUM, EIGENVALUE_LAYER_NUM_DEFAULT = "layer_num", 0
PROGRESSIVE_LAYER_DROP = "progressive_layer_drop"
PLD_ENABLED, PLD_ENABLED_DEFAULT = "enabled", False
PLD_THETA, PLD_THETA_DEFAULT = "theta", 1.0
PLD_GAMMA, PLD_GAMMA_DEFAULT = "gamma", 0.001

# ---- checkpointing ---------------------------------------------------------------------------------------------------------


class ValidationMode:
    WARN, IGNORE, FAIL = "WARN", "IGNORE", "FAIL"


CHECKPOINT = "checkpoint"
CHECKPOINT_TAG_VALIDATION, CHECKPOINT_TAG_VALIDATION_DEFAULT = "tag_validation", ValidationMode.WARN
CHECKPOINT_TAG_VALIDATION_MODES = [ValidationMode.WARN, ValidationMode.IGNORE, ValidationMode.FAIL]
LOAD_UNIVERSAL_CHECKPOINT, LOAD_UNIVERSAL_CHECKPOINT_DEFAULT = "load_universal", False
USE_NODE_LOCAL_STORAGE_CHECKPOINT, USE_NODE_LOCAL_STORAGE_CHECKPOINT_DEFAULT = "use_node_local_storage", False
CHECKPOINT_PARALLEL_WRITE = "parallel_write"
CHECKPOINT_PARALLEL_WRITE_PIPELINE_STAGE, CHECKPOINT_PARALLEL_WRITE_PIPELINE_STAGE_DEFAULT = "pipeline_stage", False

# ---- sparse attention ----------------------------------------------------------------------------------------------------
SPARSE_ATTENTION = "sparse_attention"
SPARSE_DENSE_MODE, SPARSE_FIXED_MODE, SPARSE_VARIABLE_MODE = "dense", "fixed", "variable"
SPARSE_BIGBIRD_MODE, SPARSE_BSLONGFORMER_MODE = "bigbird", "bslongformer"
SPARSE_MODE, SPARSE_MODE_DEFAULT = "mode", SPARSE_FIXED_MODE
SPARSE_BLOCK, SPARSE_BLOCK_DEFAULT = "block", 16
SPARSE_DIFFERENT_LAYOUT_PER_HEAD, SPARSE_DIFFERENT_LAYOUT_PER_HEAD_DEFAULT = "different_layout_per_head", False
SPARSE_NUM_LOCAL_BLOCKS, SPARSE_NUM_LOCAL_BLOCKS_DEFAULT = "num_local_blocks", 4
SPARSE_NUM_GLOBAL_BLOCKS, SPARSE_NUM_GLOBAL_BLOCKS_DEFAULT = "num_global_blocks", 1
SPARSE_ATTENTION_TYPE, SPARSE_ATTENTION_TYPE_DEFAULT = "attention", "bidirectional"
SPARSE_HORIZONTAL_GLOBAL_ATTENTION, SPARSE_HORIZONTAL_GLOBAL_ATTENTION_DEFAULT = "horizontal_global_attention", False
SPARSE_NUM_DIFFERENT_GLOBAL_PATTERNS, SPARSE_NUM_DIFFERENT_GLOBAL_PATTERNS_DEFAULT = "num_different_global_patterns", 1
SPARSE_NUM_RANDOM_BLOCKS, SPARSE_NUM_RANDOM_BLOCKS_DEFAULT = "num_random_blocks", 0
SPARSE_LOCAL_WINDOW_BLOCKS, SPARSE_LOCAL_WINDOW_BLOCKS_DEFAULT = "local_window_blocks", [4]
SPARSE_GLOBAL_BLOCK_INDICES, SPARSE_GLOBAL_BLOCK_INDICES_DEFAULT = "global_block_indices", [0]
SPARSE_GLOBAL_BLOCK_END_INDICES, SPARSE_GLOBAL_BLOCK_END_INDICES_DEFAULT = "global_block_end_indices", None
SPARSE_NUM_SLIDING_WINDOW_BLOCKS, SPARSE_NUM_SLIDING_WINDOW_BLOCKS_DEFAULT = "num_sliding_window_blocks", 3

# ---- routing / parallel-group attribute names --------------------------------------------------------------------------
ROUTE_TRAIN, ROUTE_EVAL, ROUTE_PREDICT, ROUTE_ENCODE = "train", "eval", "predict", "encode"
PIPE_REPLICATED = "ds_pipe_replicated"
DATA_PARALLEL_GROUP = "data_parallel_group"
GLOBAL_RANK = "global_rank"
USE_DATA_BEFORE_EXPERT_PARALLEL, USE_DATA_BEFORE_EXPERT_PARALLEL_DEFAULT = "use_data_before_expert_parallelism", False
