"""``csvMonitor`` (reference ``monitor/csv_monitor.py``); the implementation lives with the other writers in ``monitor/monitor.py``."""
from .monitor import Monitor, csvMonitor  # noqa: F401
