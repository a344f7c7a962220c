"""Kept-token schedule for random-LTD (reference ``data_routing/scheduler.py:38 RandomLTDScheduler``): the number
of tokens each wrapped layer keeps grows linearly from ``min_value`` to ``max_value`` by ``seq_per_step`` every
``require_steps`` steps."""
import math

from .. import constants as C


class BaseScheduler:

    def __init__(self):
        self.state = {}

    def _linear(self, step):
        cfg = self.state[C.RANDOM_LTD_SCHEDULE_CONFIG]
        inc, req = cfg[C.RANDOM_LTD_INCREASE_STEP], cfg[C.RANDOM_LTD_REQUIRE_STEP]
        v = self.state[C.RANDOM_LTD_MIN_VALUE] + math.floor(step / req) * inc
        return min(v, self.state[C.RANDOM_LTD_MAX_VALUE])

    def get_value(self, global_steps):
        if self.state[C.RANDOM_LTD_SCHEDULER_TYPE] == "fixed_linear":
            return self._linear(global_steps)
        raise RuntimeError("Unsupported random LTD schedule type")


class RandomLTDScheduler(BaseScheduler):

    def __init__(self, config):
        super().__init__()
        self.model_layer_num = config[C.RANDOM_LTD_TOTAL_LAYER_NUM]
        self.random_ltd_layer_num = config[C.RANDOM_LTD_LAYER_NUM]
        self.config_schedule = config[C.RANDOM_LTD_SCHEDULER]
        self.global_batch_size = config[C.RANDOM_LTD_GLOBAL_BATCH_SIZE]
        self.reset_to_init()
        if config.get(C.RANDOM_LTD_LAYER_TOKEN_LR_SCHEDULE, {}).get("enabled", False) if hasattr(C, "RANDOM_LTD_LAYER_TOKEN_LR_SCHEDULE") else False:
            raise NotImplementedError

    def reset_to_init(self):
        s = self.config_schedule
        self.state = {
            C.RANDOM_LTD_MIN_VALUE: s[C.RANDOM_LTD_MIN_VALUE], C.RANDOM_LTD_MAX_VALUE: s[C.RANDOM_LTD_MAX_VALUE],
            C.RANDOM_LTD_CURRENT_VALUE: s[C.RANDOM_LTD_MIN_VALUE], C.RANDOM_LTD_SCHEDULE_CONFIG: s[C.RANDOM_LTD_SCHEDULE_CONFIG],
            C.RANDOM_LTD_SCHEDULER_TYPE: s[C.RANDOM_LTD_SCHEDULER_TYPE], C.RANDOM_LTD_CONSUMED_LAYER_TOKENS: 0,
            C.RANDOM_LTD_CURR_STEP: -1,
        }

    def get_total_layer_tokens(self, train_iters):
        total = 0
        for step in range(train_iters):
            total += self._tokens_at(self.get_value(step))
        return total

    def _tokens_at(self, kept):
        full = self.state[C.RANDOM_LTD_MAX_VALUE]
        return self.global_batch_size * (kept * self.random_ltd_layer_num + full *
                                         (self.model_layer_num - self.random_ltd_layer_num))

    def get_current_seq(self):
        return self.state[C.RANDOM_LTD_CURRENT_VALUE]

    def set_current_seq(self, seq_length):
        self.state[C.RANDOM_LTD_CURRENT_VALUE] = seq_length

    def get_random_ltd_layer_num(self):
        return self.random_ltd_layer_num

    def get_state(self):
        return self.state

    def set_state(self, state):
        self.state = state

    def update_seq(self, global_steps):
        if self.state[C.RANDOM_LTD_CURRENT_VALUE] < self.state[C.RANDOM_LTD_MAX_VALUE]:
            self.state[C.RANDOM_LTD_CURRENT_VALUE] = self.get_value(global_steps)
        if global_steps != self.state[C.RANDOM_LTD_CURR_STEP]:
            self.state[C.RANDOM_LTD_CONSUMED_LAYER_TOKENS] += self._tokens_at(self.state[C.RANDOM_LTD_CURRENT_VALUE])
            self.state[C.RANDOM_LTD_CURR_STEP] = global_steps

    def state_dict(self):
        return {k: self.state[k] for k in (C.RANDOM_LTD_CONSUMED_LAYER_TOKENS, C.RANDOM_LTD_CURR_STEP,
                                           C.RANDOM_LTD_CURRENT_VALUE, C.RANDOM_LTD_MIN_VALUE, C.RANDOM_LTD_MAX_VALUE)}

    def load_state_dict(self, sd):
        self.state.update(sd)
