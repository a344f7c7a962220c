"""``"tensor_parallel"`` training config (reference ``runtime/tensor_parallel/config.py``)."""
from enum import Enum

from deepspeed_b200.runtime.config import TensorParallelConfig as TPTrainingConfig  # noqa: F401
from deepspeed_b200.runtime.config import TensorParallelTPConfig as TPConfig  # noqa: F401


class AUTOTP_MODE(Enum):
    TRAINING = "TRAINING"
    INFERENCE = "INFERENCE"


def get_tensor_parallel_config(ds_config):
    return TPTrainingConfig(**ds_config["tensor_parallel"]) if "tensor_parallel" in ds_config else TPTrainingConfig()
