"""Monitor config models (reference ``monitor/config.py``)."""
from deepspeed_b200.runtime.config import CometConfig, CSVConfig, MonitorConfig, TensorBoardConfig, WandbConfig  # noqa: F401

DeepSpeedMonitorConfig = MonitorConfig


def get_monitor_config(param_dict):
    return MonitorConfig({k: param_dict.get(k, {}) for k in ("tensorboard", "wandb", "csv_monitor", "comet")})
