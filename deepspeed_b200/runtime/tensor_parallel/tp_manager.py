from . import TpTrainingManager  # noqa: F401  (reference ``runtime/tensor_parallel/tp_manager.py``)
