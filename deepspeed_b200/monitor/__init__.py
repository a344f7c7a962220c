from .monitor import MonitorMaster  # noqa: F401
