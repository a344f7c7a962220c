"""``WandbMonitor`` (reference ``monitor/wandb.py``); the implementation lives with the other writers in ``monitor/monitor.py``."""
from .monitor import Monitor, WandbMonitor  # noqa: F401
