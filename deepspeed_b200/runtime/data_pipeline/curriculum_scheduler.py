"""Curriculum difficulty schedules (reference ``runtime/data_pipeline/curriculum_scheduler.py:11``).

Schedules: ``fixed_discrete`` (piecewise-constant table), ``fixed_linear`` / ``fixed_root`` (difficulty grows as
``(step/total)^(1/degree)`` between min and max, floored to a multiple of ``difficulty_step``), ``custom``
(user callback).  State is a plain dict so it round-trips through checkpoints.
"""
import math

from deepspeed_b200.utils.logging import logger
from . import constants as C

_REQUIRED = {
    C.CURRICULUM_LEARNING_SCHEDULE_FIXED_DISCRETE: (C.CURRICULUM_LEARNING_SCHEDULE_DIFFICULTY,
                                                    C.CURRICULUM_LEARNING_SCHEDULE_MAX_STEP),
    C.CURRICULUM_LEARNING_SCHEDULE_FIXED_ROOT: (C.CURRICULUM_LEARNING_SCHEDULE_TOTAL_STEP,
                                                C.CURRICULUM_LEARNING_SCHEDULE_DIFFICULTY_STEP,
                                                C.CURRICULUM_LEARNING_SCHEDULE_ROOT_DEGREE),
    C.CURRICULUM_LEARNING_SCHEDULE_FIXED_LINEAR: (C.CURRICULUM_LEARNING_SCHEDULE_TOTAL_STEP,
                                                  C.CURRICULUM_LEARNING_SCHEDULE_DIFFICULTY_STEP),
    C.CURRICULUM_LEARNING_SCHEDULE_CUSTOM: (),
}


class CurriculumScheduler:

    def __init__(self, config):
        for key in (C.CURRICULUM_LEARNING_MIN_DIFFICULTY, C.CURRICULUM_LEARNING_MAX_DIFFICULTY,
                    C.CURRICULUM_LEARNING_SCHEDULE_TYPE):
            assert key in config, f"Curriculum learning requires the config '{key}'"
        kind = config[C.CURRICULUM_LEARNING_SCHEDULE_TYPE]
        if kind not in _REQUIRED:
            raise RuntimeError("Unsupported curriculum schedule type")
        sched = config.get(C.CURRICULUM_LEARNING_SCHEDULE_CONFIG, {})
        for key in _REQUIRED[kind]:
            assert key in sched, f"Curriculum learning with {kind} schedule requires the schedule_config '{key}'"
        if kind == C.CURRICULUM_LEARNING_SCHEDULE_FIXED_DISCRETE:
            diffs, steps = sched[C.CURRICULUM_LEARNING_SCHEDULE_DIFFICULTY], sched[C.CURRICULUM_LEARNING_SCHEDULE_MAX_STEP]
            assert len(steps) > 0 and len(diffs) == len(steps) + 1, \
                "fixed_discrete needs one more difficulty than max_step entries"
        elif kind in (C.CURRICULUM_LEARNING_SCHEDULE_FIXED_ROOT, C.CURRICULUM_LEARNING_SCHEDULE_FIXED_LINEAR):
            if sched[C.CURRICULUM_LEARNING_SCHEDULE_DIFFICULTY_STEP] % 8 != 0:
                logger.warning("When using seqlen metric, the difficulty_step for curriculum learning should be a "
                               "multiple of 8 (tensor-core tile granularity). Disregard if unrelated to your metric.")
        self.state = {
            C.CURRICULUM_LEARNING_MIN_DIFFICULTY: config[C.CURRICULUM_LEARNING_MIN_DIFFICULTY],
            C.CURRICULUM_LEARNING_MAX_DIFFICULTY: config[C.CURRICULUM_LEARNING_MAX_DIFFICULTY],
            C.CURRICULUM_LEARNING_CURRENT_DIFFICULTY: config[C.CURRICULUM_LEARNING_MIN_DIFFICULTY],
            C.CURRICULUM_LEARNING_SCHEDULE_TYPE: kind,
        }
        if kind != C.CURRICULUM_LEARNING_SCHEDULE_CUSTOM:
            self.state[C.CURRICULUM_LEARNING_SCHEDULE_CONFIG] = sched
        self.custom_get_difficulty = None
        self.first_step = True

    def get_current_difficulty(self):
        return self.state[C.CURRICULUM_LEARNING_CURRENT_DIFFICULTY]

    def set_current_difficulty(self, difficulty):
        self.state[C.CURRICULUM_LEARNING_CURRENT_DIFFICULTY] = difficulty

    def set_custom_get_difficulty(self, schedule_function):
        self.custom_get_difficulty = schedule_function

    def get_state(self):
        return self.state

    def set_state(self, state):
        self.state = state

    def _table(self, step):
        sched = self.state[C.CURRICULUM_LEARNING_SCHEDULE_CONFIG]
        diffs, limits = sched[C.CURRICULUM_LEARNING_SCHEDULE_DIFFICULTY], sched[C.CURRICULUM_LEARNING_SCHEDULE_MAX_STEP]
        for d, lim in zip(diffs, limits):
            if step <= lim:
                return d
        return diffs[-1]

    def _power(self, step, degree=None):
        sched = self.state[C.CURRICULUM_LEARNING_SCHEDULE_CONFIG]
        degree = degree if degree is not None else sched[C.CURRICULUM_LEARNING_SCHEDULE_ROOT_DEGREE]
        lo, hi = self.state[C.CURRICULUM_LEARNING_MIN_DIFFICULTY], self.state[C.CURRICULUM_LEARNING_MAX_DIFFICULTY]
        frac = (float(step) / sched[C.CURRICULUM_LEARNING_SCHEDULE_TOTAL_STEP])**(1.0 / degree)
        d = math.floor(frac * (hi - lo) + lo)
        d -= d % sched[C.CURRICULUM_LEARNING_SCHEDULE_DIFFICULTY_STEP]
        return min(d, hi)

    def get_difficulty(self, global_steps):
        kind = self.state[C.CURRICULUM_LEARNING_SCHEDULE_TYPE]
        if kind == C.CURRICULUM_LEARNING_SCHEDULE_FIXED_DISCRETE:
            return self._table(global_steps)
        if kind == C.CURRICULUM_LEARNING_SCHEDULE_FIXED_LINEAR:
            return self._power(global_steps, 1)
        if kind == C.CURRICULUM_LEARNING_SCHEDULE_FIXED_ROOT:
            return self._power(global_steps)
        if kind == C.CURRICULUM_LEARNING_SCHEDULE_CUSTOM:
            return self.custom_get_difficulty(global_steps)
        raise RuntimeError("Unsupported curriculum schedule type")

    def update_difficulty(self, global_steps):
        if self.state[C.CURRICULUM_LEARNING_CURRENT_DIFFICULTY] < self.state[C.CURRICULUM_LEARNING_MAX_DIFFICULTY]:
            self.state[C.CURRICULUM_LEARNING_CURRENT_DIFFICULTY] = self.get_difficulty(global_steps)
        return self.state[C.CURRICULUM_LEARNING_CURRENT_DIFFICULTY]
