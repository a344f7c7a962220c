"""Learning-rate schedules: LRRangeTest, OneCycle, WarmupLR, WarmupDecayLR, WarmupCosineLR.

Parity target: reference ``runtime/lr_schedules.py`` (``:273, :371, :633, :723, :774``).  Same
constructor arguments and the same closed-form curves; written as small stateless ``_lr_at(step)``
functions under a common stepper so ``state_dict`` is just the step counter.
"""
import argparse
import math
from typing import List, Union

LR_SCHEDULE = "lr_schedule"
LR_RANGE_TEST = "LRRangeTest"
ONE_CYCLE = "OneCycle"
WARMUP_LR = "WarmupLR"
WARMUP_DECAY_LR = "WarmupDecayLR"
WARMUP_COSINE_LR = "WarmupCosineLR"
VALID_LR_SCHEDULES = [LR_RANGE_TEST, ONE_CYCLE, WARMUP_LR, WARMUP_DECAY_LR, WARMUP_COSINE_LR]

WARMUP_LOG_RATE = "log"
WARMUP_LINEAR_RATE = "linear"


def add_tuning_arguments(parser: argparse.ArgumentParser):
    g = parser.add_argument_group("Convergence Tuning", "Convergence tuning configurations")
    g.add_argument("--lr_schedule", type=str, default=None, help="LR schedule for training.")
    g.add_argument("--lr_range_test_min_lr", type=float, default=0.001)
    g.add_argument("--lr_range_test_step_rate", type=float, default=1.0)
    g.add_argument("--lr_range_test_step_size", type=int, default=1000)
    g.add_argument("--lr_range_test_staircase", type=bool, default=False)
    g.add_argument("--cycle_first_step_size", type=int, default=1000)
    g.add_argument("--cycle_first_stair_count", type=int, default=-1)
    g.add_argument("--cycle_second_step_size", type=int, default=-1)
    g.add_argument("--cycle_second_stair_count", type=int, default=-1)
    g.add_argument("--decay_step_size", type=int, default=1000)
    g.add_argument("--cycle_min_lr", type=float, default=0.01)
    g.add_argument("--cycle_max_lr", type=float, default=0.1)
    g.add_argument("--decay_lr_rate", type=float, default=0.0)
    g.add_argument("--cycle_momentum", default=False, action="store_true")
    g.add_argument("--cycle_min_mom", type=float, default=0.8)
    g.add_argument("--cycle_max_mom", type=float, default=0.9)
    g.add_argument("--decay_mom_rate", type=float, default=0.0)
    g.add_argument("--warmup_min_lr", type=float, default=0)
    g.add_argument("--warmup_max_lr", type=float, default=0.001)
    g.add_argument("--warmup_num_steps", type=int, default=1000)
    g.add_argument("--warmup_type", type=str, default=WARMUP_LOG_RATE)
    return parser


def get_torch_optimizer(optimizer):
    if hasattr(optimizer, "param_groups") and not hasattr(optimizer, "optimizer"):
        return optimizer
    if hasattr(optimizer, "optimizer"):
        return optimizer.optimizer
    raise TypeError(f"{type(optimizer).__name__} is not a torch-style optimizer")


def _as_list(v, n, name):
    if isinstance(v, (list, tuple)):
        if len(v) != n:
            raise ValueError(f"expected {n} values for {name}, got {len(v)}")
        return list(v)
    return [v] * n


class _Schedule:
    """Common stepper: subclasses implement ``get_lr()`` from ``self.last_batch_iteration``."""

    def __init__(self, optimizer, last_batch_iteration=-1):
        self.optimizer = get_torch_optimizer(optimizer)
        self.last_batch_iteration = last_batch_iteration
        self._last_lr = [g["lr"] for g in self.optimizer.param_groups]

    def get_lr(self) -> List[float]:
        raise NotImplementedError

    def get_last_lr(self):
        return self._last_lr

    def _apply(self, lrs):
        for g, lr in zip(self.optimizer.param_groups, lrs):
            g["lr"] = lr
        self._last_lr = list(lrs)

    def step(self, last_batch_iteration=None):
        if last_batch_iteration is None:
            last_batch_iteration = self.last_batch_iteration + 1
        self.last_batch_iteration = last_batch_iteration
        self._apply(self.get_lr())

    def state_dict(self):
        return {"last_batch_iteration": self.last_batch_iteration}

    def load_state_dict(self, sd):
        self.last_batch_iteration = sd["last_batch_iteration"]


class LRRangeTest(_Schedule):
    """lr = min_lr * (1 + step_rate * interval)  with interval = step/step_size (floored if staircase)."""

    def __init__(self, optimizer, lr_range_test_min_lr: Union[float, list] = 1e-3, lr_range_test_step_size=2000,
                 lr_range_test_step_rate=1.0, lr_range_test_staircase=False, last_batch_iteration=-1):
        super().__init__(optimizer, last_batch_iteration)
        n = len(self.optimizer.param_groups)
        self.min_lr = _as_list(lr_range_test_min_lr, n, "lr_range_test_min_lr")
        self.step_size = lr_range_test_step_size
        self.step_rate = lr_range_test_step_rate
        self.staircase = lr_range_test_staircase
        if last_batch_iteration == -1:
            self._apply(self.min_lr)

    def _interval(self):
        x = float(self.last_batch_iteration + 1) / self.step_size
        return math.floor(x) if self.staircase else x

    def get_lr(self):
        f = 1 + self.step_rate * self._interval()
        return [m * f for m in self.min_lr]


class OneCycle(_Schedule):
    """Triangular cycle min->max->min then decay; optional inverse momentum cycle."""

    def __init__(self, optimizer, cycle_min_lr, cycle_max_lr, decay_lr_rate=0.0, cycle_first_step_size=2000,
                 cycle_second_step_size=None, cycle_first_stair_count=0, cycle_second_stair_count=None,
                 decay_step_size=0, cycle_momentum=True, cycle_min_mom=0.8, cycle_max_mom=0.9, decay_mom_rate=0.0,
                 last_batch_iteration=-1):
        super().__init__(optimizer, last_batch_iteration)
        n = len(self.optimizer.param_groups)
        first = float(cycle_first_step_size)
        second = float(cycle_second_step_size) if cycle_second_step_size is not None else first
        self.total_size = first + second
        self.step_ratio = first / self.total_size
        self.first_stair_count = cycle_first_stair_count
        self.second_stair_count = cycle_first_stair_count if cycle_second_stair_count is None else cycle_second_stair_count
        self.decay_step_size = decay_step_size
        self.min_lrs = _as_list(cycle_min_lr, n, "cycle_min_lr")
        self.max_lrs = _as_list(cycle_max_lr, n, "cycle_max_lr")
        self.decay_lr_rate = decay_lr_rate
        self.cycle_momentum = cycle_momentum
        if cycle_momentum:
            g0 = self.optimizer.param_groups[0]
            if "betas" not in g0 and "momentum" not in g0:
                self.cycle_momentum = False
            else:
                self.min_moms = _as_list(cycle_min_mom, n, "cycle_min_mom")
                self.max_moms = _as_list(cycle_max_mom, n, "cycle_max_mom")
                self.decay_mom_rate = decay_mom_rate
        if last_batch_iteration == -1:
            self._apply(self.min_lrs)
            if self.cycle_momentum:
                self._apply_mom(self.max_moms)

    def _apply_mom(self, moms):
        for g, m in zip(self.optimizer.param_groups, moms):
            m = m[0] if isinstance(m, (tuple, list)) else m
            if "betas" in g:
                g["betas"] = (m, g["betas"][1])
            else:
                g["momentum"] = m

    def _mom_out(self, values):
        """Per group: ``(beta1, beta2)`` for Adam-style optimizers, the momentum scalar otherwise (reference
        ``lr_schedules.py:592 get_mom`` returns the ``betas`` pairs)."""
        out = []
        for g, m in zip(self.optimizer.param_groups, values):
            out.append((m, g["betas"][1]) if "betas" in g else m)
        return out

    def _scale(self):
        it = self.last_batch_iteration + 1
        cycle = math.floor(1 + it / self.total_size)
        x = 1.0 + it / self.total_size - cycle
        return x / self.step_ratio if x <= self.step_ratio else (x - 1) / (self.step_ratio - 1)

    def get_lr(self):
        it = self.last_batch_iteration + 1
        if it < self.total_size:
            s = self._scale()
            return [lo + (hi - lo) * s for lo, hi in zip(self.min_lrs, self.max_lrs)]
        decay_it = it - self.total_size + 1
        interval = decay_it / self.decay_step_size if self.decay_step_size else 0.0
        f = 1 + self.decay_lr_rate * interval
        return [lo / f for lo in self.min_lrs]

    def get_mom(self):
        if not self.cycle_momentum:
            return None
        it = self.last_batch_iteration + 1
        if it < self.total_size:
            s = self._scale()
            return self._mom_out([hi - (hi - lo) * s for lo, hi in zip(self.min_moms, self.max_moms)])
        decay_it = it - self.total_size + 1
        interval = decay_it / self.decay_step_size if self.decay_step_size else 0.0
        f = 1 + self.decay_mom_rate * interval
        return self._mom_out([hi * f for hi in self.max_moms])

    def step(self, batch_iteration=None):
        super().step(batch_iteration)
        if self.cycle_momentum:
            self._apply_mom(self.get_mom())


class WarmupLR(_Schedule):
    """Warm up min->max over ``warmup_num_steps`` (log or linear), then hold."""

    def __init__(self, optimizer, warmup_min_lr: Union[float, list] = 0.0, warmup_max_lr: Union[float, list] = 0.001,
                 warmup_num_steps: int = 1000, warmup_type: str = WARMUP_LOG_RATE, last_batch_iteration: int = -1):
        super().__init__(optimizer, last_batch_iteration)
        n = len(self.optimizer.param_groups)
        self.min_lrs = _as_list(warmup_min_lr, n, "warmup_min_lr")
        self.max_lrs = _as_list(warmup_max_lr, n, "warmup_max_lr")
        self.delta_lrs = [b - s for b, s in zip(self.max_lrs, self.min_lrs)]
        if warmup_type not in (WARMUP_LOG_RATE, WARMUP_LINEAR_RATE):
            from deepspeed_b200.utils.logging import logger
            logger.warning(f"Using unknown warmup_type: {warmup_type}. The increasing function is set to default (log)")
            warmup_type = WARMUP_LOG_RATE
        self.warmup_type = warmup_type
        self.warmup_num_steps = max(2, warmup_num_steps)
        self.inverse_log_warm_up = 1.0 / math.log(self.warmup_num_steps)
        if last_batch_iteration == -1:
            self._apply(self.get_lr())

    def _gamma(self):
        it = self.last_batch_iteration
        if it < self.warmup_num_steps:
            if self.warmup_type == WARMUP_LOG_RATE:
                return self.inverse_log_warm_up * math.log(it + 1)
            return it / self.warmup_num_steps
        return 1.0

    def get_lr(self):
        if self.last_batch_iteration < 0:
            return list(self.min_lrs)
        g = self._gamma()
        return [lo + d * g for lo, d in zip(self.min_lrs, self.delta_lrs)]


class WarmupDecayLR(WarmupLR):
    """WarmupLR followed by linear decay to ``warmup_min_lr`` at ``total_num_steps``."""

    def __init__(self, optimizer, total_num_steps: int, warmup_min_lr=0.0, warmup_max_lr=0.001, warmup_num_steps=1000,
                 warmup_type=WARMUP_LOG_RATE, last_batch_iteration=-1):
        self.total_num_steps = total_num_steps
        super().__init__(optimizer, warmup_min_lr, warmup_max_lr, warmup_num_steps, warmup_type, last_batch_iteration)
        if self.total_num_steps < self.warmup_num_steps:
            from deepspeed_b200.utils.logging import logger
            logger.warning(f"total_num_steps {total_num_steps} is less than warmup_num_steps {warmup_num_steps}")

    def _gamma(self):
        it = self.last_batch_iteration
        if it < self.warmup_num_steps:
            return super()._gamma()
        return max(0.0, float(self.total_num_steps - it) / float(max(1.0, self.total_num_steps - self.warmup_num_steps)))


class WarmupCosineLR(_Schedule):
    """Linear/log warm-up of the *ratio* then cosine decay to ``cos_min_ratio`` (ratios scale the
    optimizer's initial lr, as in the reference :774)."""

    def __init__(self, optimizer, total_num_steps: int, warmup_min_ratio: float = 0.0, warmup_num_steps: int = 1000,
                 cos_min_ratio: float = 0.0001, warmup_type: str = WARMUP_LOG_RATE, last_batch_iteration: int = -1):
        super().__init__(optimizer, last_batch_iteration)
        self.total_num_steps = total_num_steps
        self.cos_min_ratio = cos_min_ratio
        self.warmup_type = warmup_type if warmup_type in (WARMUP_LOG_RATE, WARMUP_LINEAR_RATE) else WARMUP_LOG_RATE
        self.warmup_min_ratio = warmup_min_ratio
        self.warmup_num_steps = max(2, warmup_num_steps)
        self.inverse_log_warm_up = 1.0 / math.log(self.warmup_num_steps)
        self.org_lrs = [g["lr"] for g in self.optimizer.param_groups]
        if last_batch_iteration == -1:
            self._apply(self.get_lr())

    def get_lr_ratio(self):
        it = self.last_batch_iteration
        if it < 0:
            return 0.0
        if it < self.warmup_num_steps:
            if self.warmup_type == WARMUP_LOG_RATE:
                r = self.inverse_log_warm_up * math.log(it + 1)
            else:
                r = it / self.warmup_num_steps
            return self.warmup_min_ratio + (1.0 - self.warmup_min_ratio) * r
        real_last = it - self.warmup_num_steps + 1
        real_total = max(1, self.total_num_steps - self.warmup_num_steps)
        cos = 0.5 * (1 + math.cos(math.pi * min(real_last / real_total, 1.0)))
        return max(0.0, self.cos_min_ratio + (1.0 - self.cos_min_ratio) * cos)

    def get_lr(self):
        if self.last_batch_iteration < 0:
            return [0.0 for _ in self.org_lrs]
        r = self.get_lr_ratio()
        return [lr * r for lr in self.org_lrs]


_SCHEDULES = {
    LR_RANGE_TEST: LRRangeTest,
    ONE_CYCLE: OneCycle,
    WARMUP_LR: WarmupLR,
    WARMUP_DECAY_LR: WarmupDecayLR,
    WARMUP_COSINE_LR: WarmupCosineLR,
}


def get_lr_schedule_class(name):
    return _SCHEDULES.get(name)


def get_config_from_args(args):
    if not hasattr(args, LR_SCHEDULE) or args.lr_schedule is None:
        return None, "--lr_schedule not specified on command line"
    if args.lr_schedule not in VALID_LR_SCHEDULES:
        return None, f"{args.lr_schedule} is not supported LR schedule"
    cfg = {"type": args.lr_schedule, "params": {}}
    keys = {
        LR_RANGE_TEST: ["lr_range_test_min_lr", "lr_range_test_step_rate", "lr_range_test_step_size",
                        "lr_range_test_staircase"],
        ONE_CYCLE: ["cycle_first_step_size", "cycle_first_stair_count", "cycle_second_step_size",
                    "cycle_second_stair_count", "decay_step_size", "cycle_min_lr", "cycle_max_lr", "decay_lr_rate",
                    "cycle_min_mom", "cycle_max_mom", "decay_mom_rate"],
        WARMUP_LR: ["warmup_min_lr", "warmup_max_lr", "warmup_num_steps", "warmup_type"],
        WARMUP_DECAY_LR: ["warmup_min_lr", "warmup_max_lr", "warmup_num_steps", "warmup_type"],
        WARMUP_COSINE_LR: ["warmup_num_steps", "warmup_type"],
    }[args.lr_schedule]
    for k in keys:
        if hasattr(args, k):
            cfg["params"][k] = getattr(args, k)
    return cfg, None


# ---- parameter-name constants + CLI override helpers (reference ``lr_schedules.py:22-58, 124-262``) -------------------------
LR_RANGE_TEST_MIN_LR, LR_RANGE_TEST_STEP_RATE = "lr_range_test_min_lr", "lr_range_test_step_rate"
LR_RANGE_TEST_STEP_SIZE, LR_RANGE_TEST_STAIRCASE = "lr_range_test_step_size", "lr_range_test_staircase"
EDGE_VALUE, MID_VALUE = "edge_value", "mid_value"
CYCLE_FIRST_STEP_SIZE, CYCLE_FIRST_STAIR_COUNT = "cycle_first_step_size", "cycle_first_stair_count"
CYCLE_SECOND_STEP_SIZE, CYCLE_SECOND_STAIR_COUNT = "cycle_second_step_size", "cycle_second_stair_count"
DECAY_STEP_SIZE = "decay_step_size"
CYCLE_MIN_LR, CYCLE_MAX_LR, DECAY_LR_RATE = "cycle_min_lr", "cycle_max_lr", "decay_lr_rate"
CYCLE_MIN_MOM, CYCLE_MAX_MOM, DECAY_MOM_RATE = "cycle_min_mom", "cycle_max_mom", "decay_mom_rate"
WARMUP_MIN_LR, WARMUP_MAX_LR, WARMUP_NUM_STEPS, WARMUP_TYPE = "warmup_min_lr", "warmup_max_lr", "warmup_num_steps", "warmup_type"
WARMUP_MIN_RATIO, COS_MIN_RATIO, TOTAL_NUM_STEPS = "warmup_min_ratio", "cos_min_ratio", "total_num_steps"

_ARG_KEYS = {
    LR_RANGE_TEST: (LR_RANGE_TEST_MIN_LR, LR_RANGE_TEST_STEP_RATE, LR_RANGE_TEST_STEP_SIZE, LR_RANGE_TEST_STAIRCASE),
    ONE_CYCLE: (CYCLE_FIRST_STEP_SIZE, CYCLE_FIRST_STAIR_COUNT, CYCLE_SECOND_STEP_SIZE, CYCLE_SECOND_STAIR_COUNT, DECAY_STEP_SIZE,
                CYCLE_MIN_LR, CYCLE_MAX_LR, DECAY_LR_RATE, CYCLE_MIN_MOM, CYCLE_MAX_MOM, DECAY_MOM_RATE),
    WARMUP_LR: (WARMUP_MIN_LR, WARMUP_MAX_LR, WARMUP_NUM_STEPS, WARMUP_TYPE),
}


def parse_arguments():
    """Parse the convergence-tuning flags from ``sys.argv`` (unknown flags are returned separately)."""
    parser = add_tuning_arguments(argparse.ArgumentParser())
    return parser.parse_known_args()


def _override(args, params, keys):
    for k in keys:
        v = getattr(args, k, None)
        if v is not None:
            params[k] = v


def override_lr_range_test_params(args, params):
    _override(args, params, _ARG_KEYS[LR_RANGE_TEST])


def override_1cycle_params(args, params):
    _override(args, params, _ARG_KEYS[ONE_CYCLE])


def override_warmupLR_params(args, params):
    _override(args, params, _ARG_KEYS[WARMUP_LR])


def override_params(args, params):
    """Command-line values win over the config file for every schedule family."""
    for keys in _ARG_KEYS.values():
        _override(args, params, keys)


def get_lr_from_config(config):
    """-> (peak lr implied by a scheduler config block, error message)."""
    if "type" not in config:
        return None, "LR schedule type not defined in config"
    if "params" not in config:
        return None, "LR schedule params not defined in config"
    kind, params = config["type"], config["params"]
    if kind not in VALID_LR_SCHEDULES:
        return None, f"{kind} is not a valid LR schedule"
    key = {LR_RANGE_TEST: LR_RANGE_TEST_MIN_LR, ONE_CYCLE: CYCLE_MAX_LR}.get(kind, WARMUP_MAX_LR)
    return params[key], ""


def update_lr(param_groups, lrs):
    for group, lr in zip(param_groups, lrs):
        group["lr"] = lr
    return [g["lr"] for g in param_groups]
