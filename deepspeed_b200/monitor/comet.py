"""``CometMonitor`` (reference ``monitor/comet.py``); the implementation lives with the other writers in ``monitor/monitor.py``."""
from .monitor import Monitor, CometMonitor  # noqa: F401
