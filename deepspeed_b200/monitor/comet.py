"""``CometMonitor`` (reference ``monitor/comet.py``); the implementation lives with the other writers in ``monitor/monitor.py``."""
from .monitor import Monitor, CometMonitor  # noqa: F401


class EventsLogScheduler:
    """Rate limiter: an event name is logged at most once every ``samples_log_interval`` samples (reference ``comet.py:74``)."""

    def __init__(self, samples_log_interval: int):
        self._interval = samples_log_interval
        self._last = {}

    def needs_logging(self, name: str, current_sample: int) -> bool:
        prev = self._last.get(name)
        if prev is None or current_sample - prev >= self._interval:
            self._last[name] = current_sample
            return True
        return False
