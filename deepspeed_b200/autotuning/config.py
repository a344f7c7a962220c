"""``autotuning`` config block (reference ``autotuning/config.py``)."""
from typing import Dict, Optional

from deepspeed_b200.runtime.config_utils import DeepSpeedConfigModel


class DeepSpeedAutotuningConfig(DeepSpeedConfigModel):
    enabled: bool = False
    fast: bool = True
    results_dir: str = "autotuning_results"
    exps_dir: str = "autotuning_exps"
    overwrite: bool = True
    start_profile_step: int = 3
    end_profile_step: int = 5
    metric_path: Optional[str] = None
    tuner_type: str = "gridsearch"
    tuner_early_stopping: int = 5
    tuner_num_trials: int = 50
    arg_mappings: Optional[Dict[str, str]] = None
    model_info: Optional[Dict] = None
    model_info_path: Optional[str] = None
    mp_size: int = 1
    metric: str = "throughput"
    max_train_batch_size: Optional[int] = None
    min_train_batch_size: int = 1
    max_train_micro_batch_size_per_gpu: int = 1024
    min_train_micro_batch_size_per_gpu: int = 1
    num_tuning_micro_batch_sizes: int = 3
    zero_stages: Optional[list] = None


def get_autotuning_config(param_dict):
    return DeepSpeedAutotuningConfig(**(param_dict.get("autotuning") or {}))


MODEL_INFO_KEY_DEFAULT_DICT = {"profile": False, "num_params": None, "hidden_size": None, "num_layers": None}


def get_model_info_config(param_dict):
    """The ``autotuning.model_info`` section with defaults filled in, or ``None`` when absent."""
    sec = param_dict.get("model_info")
    if sec is None:
        return None
    return {k: sec.get(k, d) for k, d in MODEL_INFO_KEY_DEFAULT_DICT.items()}


def get_default_model_info_config():
    return dict(MODEL_INFO_KEY_DEFAULT_DICT)
