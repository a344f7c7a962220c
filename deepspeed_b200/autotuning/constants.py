AUTOTUNING = "autotuning"
AUTOTUNING_ENABLED = "enabled"
AUTOTUNING_METRIC_LATENCY = "latency"
AUTOTUNING_METRIC_THROUGHPUT = "throughput"
AUTOTUNING_METRIC_FLOPS = "flops"
AUTOTUNING_TUNER_GRIDSEARCH = "gridsearch"
AUTOTUNING_TUNER_RANDOM = "random"
AUTOTUNING_TUNER_MODELBASED = "model_based"
DEFAULT_HF_CONFIG = {"train_batch_size": "auto", "train_micro_batch_size_per_gpu": "auto", "gradient_accumulation_steps": "auto"}
DEFAULT_MIN_MEM_CONFIG = {"train_micro_batch_size_per_gpu": 1, "zero_optimization": {"stage": 3},
                          "memory_break_down": False}
# per-stage tuning spaces (lists = alternatives)
DEFAULT_TUNING_SPACE_ZERO_0 = {"zero_optimization": {"stage": 0}}
DEFAULT_TUNING_SPACE_ZERO_1 = {"zero_optimization": {"stage": 1, "reduce_bucket_size": [5e7, 5e8, 1e9],
                                                     "allgather_bucket_size": [5e7, 5e8, 1e9]}}
DEFAULT_TUNING_SPACE_ZERO_2 = {"zero_optimization": {"stage": 2, "overlap_comm": [True, False],
                                                     "reduce_scatter": [False, True], "reduce_bucket_size": [5e7, 5e8, 1e9],
                                                     "allgather_bucket_size": [5e7, 5e8, 1e9],
                                                     "contiguous_gradients": [False, True]}}
DEFAULT_TUNING_SPACE_ZERO_3 = {"zero_optimization": {"stage": 3, "overlap_comm": [True, False],
                                                     "b200_unit_prefetch": [1, 2, 4],
                                                     "b200_fused_collectives": [None, False],
                                                     "stage3_param_persistence_threshold": [1e4, 1e5, 1e6]}}
GLOBAL_TUNING_SPACE = "global"
TUNING_MICRO_BATCH_SIZE_PREFIX = "z"
MODEL_INFO = "model_info"
MODEL_INFO_PROFILE = "profile"
MODEL_INFO_NUM_PARAMS = "num_params"
MODEL_INFO_HIDDEN_SIZE = "hidden_size"
MODEL_INFO_NUM_LAYERS = "num_layers"
