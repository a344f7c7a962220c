"""``TensorBoardMonitor`` (reference ``monitor/tensorboard.py``); the implementation lives with the other writers in ``monitor/monitor.py``."""
from .monitor import Monitor, TensorBoardMonitor  # noqa: F401
