"""TP training config/manager (reference ``runtime/tensor_parallel/{config,tp_manager}.py``)."""
from deepspeed_b200.runtime.config import TensorParallelConfig as TPTrainingConfig  # noqa: F401
from deepspeed_b200.module_inject.auto_tp import tp_model_init  # noqa: F401


class TpTrainingManager:

    def __init__(self, model, tp_size, dtype):
        self.module = tp_model_init(model, tp_size, dtype)
