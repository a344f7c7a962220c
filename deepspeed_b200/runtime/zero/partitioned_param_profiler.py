"""Counters for ZeRO-3 parameter traffic (reference ``runtime/zero/partitioned_param_profiler.py``): how many unit
fetches were already prefetched (hits) vs issued on demand (misses) and how many elements moved.  Enabled with
``DSB200_ZERO_PROFILE=1``; ``log_events()`` prints and resets."""
import os
from collections import defaultdict

from deepspeed_b200.utils.logging import log_dist


class PartitionedParameterProfiler:

    @staticmethod
    def enabled():
        return os.environ.get("DSB200_ZERO_PROFILE", "0") == "1"

    def __init__(self, timers=None):
        self.timers = timers
        self.event_counters = defaultdict(lambda: [0, 0])

    def reset_events(self):
        self.event_counters.clear()

    def start_event(self, name):
        if self.timers is not None:
            self.timers(name).start()

    def stop_event(self, name, num_elem=0):
        if self.timers is not None:
            self.timers(name).stop()
        self.count(name, num_elem)

    def count(self, name, num_elem=0):
        c = self.event_counters[name]
        c[0] += 1
        c[1] += int(num_elem)

    def log_events(self):
        for name, (n, elems) in sorted(self.event_counters.items()):
            log_dist(f"zero3 profile: {name}: count={n} numel={elems}", ranks=[0])
        self.reset_events()
