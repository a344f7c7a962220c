"""Phase timers and throughput meter.

Parity target: reference ``utils/timer.py`` (``SynchronizedWallClockTimer :44``,
``ThroughputTimer :199``).  Timers use CUDA events on the current stream; elapsed values are
resolved only when read so the training loop never synchronises for timing.
"""
import time

import torch

from deepspeed_b200.utils.logging import log_dist

FORWARD_MICRO_TIMER = "fwd_microstep"
FORWARD_GLOBAL_TIMER = "fwd"
BACKWARD_MICRO_TIMER = "bwd_microstep"
BACKWARD_GLOBAL_TIMER = "bwd"
BACKWARD_INNER_MICRO_TIMER = "bwd_inner_microstep"
BACKWARD_INNER_GLOBAL_TIMER = "bwd_inner"
BACKWARD_REDUCE_MICRO_TIMER = "bwd_allreduce_microstep"
BACKWARD_REDUCE_GLOBAL_TIMER = "bwd_allreduce"
STEP_MICRO_TIMER = "step_microstep"
STEP_GLOBAL_TIMER = "step"
TIME_EPSILON = 1e-6


def _use_cuda_events():
    return torch.cuda.is_available()


class _EventTimer:
    """One named timer: list of (start, end) event pairs, lazily reduced to milliseconds."""

    def __init__(self, name):
        self.name_ = name
        self.started_ = False
        self._pairs = []
        self._start = None
        self._records_ms = []

    def start(self):
        assert not self.started_, f"{self.name_} timer has already been started"
        if _use_cuda_events():
            self._start = torch.cuda.Event(enable_timing=True)
            self._start.record()
        else:
            self._start = time.perf_counter()
        self.started_ = True

    def stop(self, reset=False, record=False):
        assert self.started_, f"{self.name_} timer is not started"
        if _use_cuda_events():
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            self._pairs.append((self._start, end))
        else:
            self._records_ms.append((time.perf_counter() - self._start) * 1e3)
        self.started_ = False

    def _resolve(self):
        for s, e in self._pairs:
            e.synchronize()
            self._records_ms.append(s.elapsed_time(e))
        self._pairs.clear()

    def elapsed(self, reset=True):
        was_started = self.started_
        if was_started:
            self.stop()
        self._resolve()
        total = sum(self._records_ms)
        if reset:
            self.reset()
        if was_started:
            self.start()
        return total

    def mean(self):
        self._resolve()
        if not self._records_ms:
            return 0.0
        return sum(self._records_ms) / len(self._records_ms)

    def reset(self):
        self.started_ = False
        self._pairs.clear()
        self._records_ms.clear()


class SynchronizedWallClockTimer:

    Timer = _EventTimer

    def __init__(self):
        self.timers = {}

    def get_timers(self):
        return self.timers

    def __call__(self, name):
        if name not in self.timers:
            self.timers[name] = _EventTimer(name)
        return self.timers[name]

    def has(self, name):
        return name in self.timers

    @staticmethod
    def memory_usage():
        if not torch.cuda.is_available():
            return ""
        gb = 1024**3
        return (f" | mem alloc {torch.cuda.memory_allocated() / gb:.2f} GB (max "
                f"{torch.cuda.max_memory_allocated() / gb:.2f}) reserved {torch.cuda.memory_reserved() / gb:.2f} GB")

    def log(self, names, normalizer=1.0, reset=True, memory_breakdown=False, ranks=None):
        assert normalizer > 0.0
        parts = []
        for name in names:
            if name in self.timers:
                parts.append(f"{name}: {self.timers[name].elapsed(reset=reset) / normalizer:.2f}")
        msg = "time (ms) | " + " | ".join(parts)
        if memory_breakdown:
            msg += self.memory_usage()
        log_dist(msg, ranks=ranks or [0])

    def get_mean(self, names, normalizer=1.0, reset=True):
        out = {}
        for name in names:
            if name in self.timers:
                out[name] = self.timers[name].mean() / normalizer
                if reset:
                    self.timers[name].reset()
        return out


class NoopTimer:

    class Timer:

        def start(self):
            pass

        def reset(self):
            pass

        def stop(self, **kw):
            pass

        def elapsed(self, **kw):
            return 0

        def mean(self):
            return 0

    def __init__(self):
        self.timer = self.Timer()

    def __call__(self, name):
        return self.timer

    def get_timers(self):
        return {}

    def has(self, name):
        return False

    def log(self, *a, **k):
        pass

    def get_mean(self, *a, **k):
        return {}


class ThroughputTimer:
    """samples/s meter (reference: utils/timer.py:199).

    ``start()``/``stop(global_step=True)`` bracket one micro step; the first ``start_step``
    global steps are excluded as warm-up.  When ``synchronized`` the window edges call
    ``torch.cuda.synchronize`` exactly like the reference; otherwise CUDA events are used.
    """

    def __init__(self, config=None, batch_size=1, start_step=2, steps_per_output=None, monitor_memory=False,
                 logging_fn=None, synchronized=True, enabled=True):
        if config is not None:
            enabled = getattr(config, "enabled", enabled)
            synchronized = getattr(config, "synchronized", synchronized)
        self.enabled = enabled
        self.synchronized = synchronized
        self.batch_size = max(1, batch_size or 1)
        self.start_step = start_step
        self.steps_per_output = steps_per_output
        self.monitor_memory = monitor_memory
        self.logging = logging_fn or (lambda m: log_dist(m, ranks=[0]))
        self.initialized = False
        self.started = False
        self.start_time = 0.0
        self.end_time = 0.0
        self.epoch_count = 0
        self.micro_step_count = 0
        self.global_step_count = 0
        self.total_elapsed_time = 0.0
        self.step_elapsed_time = 0.0

    def update_epoch_count(self):
        self.epoch_count += 1
        self.micro_step_count = 0

    def _init_timer(self):
        self.initialized = True

    def _sync(self):
        if self.synchronized and torch.cuda.is_available():
            torch.cuda.synchronize()

    def start(self):
        if not self.enabled:
            return
        self._init_timer()
        self.started = True
        if self.global_step_count >= self.start_step:
            self._sync()
            self.start_time = time.perf_counter()

    def stop(self, global_step=False, report_speed=True):
        if not self.enabled or not self.started:
            return
        self.started = False
        self.micro_step_count += 1
        if global_step:
            self.global_step_count += 1
        if self.start_time > 0:
            self._sync()
            self.end_time = time.perf_counter()
            duration = self.end_time - self.start_time
            self.total_elapsed_time += duration
            self.step_elapsed_time += duration
            if global_step:
                if report_speed and self.steps_per_output and self.global_step_count % self.steps_per_output == 0:
                    self.logging(
                        f"epoch={self.epoch_count}/micro_step={self.micro_step_count}/global_step={self.global_step_count}, "
                        f"RunningAvgSamplesPerSec={self.avg_samples_per_sec():.3f}, "
                        f"CurrSamplesPerSec={self.batch_size / max(self.step_elapsed_time, TIME_EPSILON):.3f}" +
                        (SynchronizedWallClockTimer.memory_usage() if self.monitor_memory else ""))
                self.step_elapsed_time = 0.0

    def avg_samples_per_sec(self):
        if self.global_step_count > self.start_step and self.total_elapsed_time > 0:
            steps = self.global_step_count - self.start_step
            return self.batch_size / (self.total_elapsed_time / steps)
        return float("-inf")


def trim_mean(data, trim_percent):
    """Mean after dropping ``trim_percent`` of samples from each tail."""
    assert 0.0 <= trim_percent <= 1.0
    n = len(data)
    if n == 0:
        return 0
    data = sorted(data)
    k = int(round(n * trim_percent))
    kept = data[k:n - k] or data
    return sum(kept) / len(kept)


CudaEventTimer = _EventTimer  # reference name


def mean(values):
    return sum(values) / len(values) if values else 0.0
