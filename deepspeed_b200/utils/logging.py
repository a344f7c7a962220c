"""Rank-aware logging helpers.

Parity target: reference ``deepspeed/utils/logging.py`` (``logger``, ``log_dist``,
``print_json_dist``, ``warning_once``).  Implementation is independent: a single
module-level logger whose handler tags every record with the rank read lazily from
the environment / torch.distributed.
"""
import functools
import json
import logging
import os
import sys

_LEVELS = {
    "debug": logging.DEBUG,
    "info": logging.INFO,
    "warning": logging.WARNING,
    "error": logging.ERROR,
    "critical": logging.CRITICAL,
}


def _current_rank() -> int:
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0"))


class _RankFilter(logging.Filter):

    def filter(self, record):
        record.rank = _current_rank()
        return True


def _build_logger(name="deepspeed_b200", level=logging.INFO):
    lg = logging.getLogger(name)
    if getattr(lg, "_dsb200_configured", False):
        return lg
    lg.setLevel(_LEVELS.get(os.environ.get("DSB200_LOG_LEVEL", "").lower(), level))
    lg.propagate = False
    h = logging.StreamHandler(stream=sys.stdout)
    h.setFormatter(
        logging.Formatter("[%(asctime)s] [%(levelname)s] [rank %(rank)s] [%(filename)s:%(lineno)d] %(message)s"))
    h.addFilter(_RankFilter())
    lg.addHandler(h)
    lg._dsb200_configured = True
    return lg


logger = _build_logger()


class LoggerFactory:
    """Reference-compatible factory (``utils/logging.py:LoggerFactory``)."""

    @staticmethod
    def create_logger(name=None, level=logging.INFO):
        if name is None:
            raise ValueError("name for logger cannot be None")
        return _build_logger(name, level)

    @staticmethod
    def create_warning_filter(logger):
        """Logging filter that warns ONCE if logging happens while a graph is being captured / traced (host-side logging
        inside a captured region is a silent no-op on replay)."""
        warned = False

        def warn_once(record):
            nonlocal warned
            if not warned and _is_capturing():
                warned = True
                logger.warning("logging inside a captured / traced region: messages will not repeat on replay "
                               "(set DISABLE_LOGS_WHILE_COMPILING=1 to silence)")
            return True

        return warn_once

    @staticmethod
    def logging_decorator(func):
        """Skip the wrapped logging call entirely while capturing / tracing."""
        import functools

        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            if _is_capturing():
                return None
            return func(*args, **kwargs)

        return wrapper


def _is_capturing():
    try:
        import torch
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            return True
        return bool(getattr(torch.compiler, "is_compiling", lambda: False)())
    except Exception:
        return False


def should_log(ranks=None) -> bool:
    """True when this process' rank is in ``ranks`` (``None``/``[-1]`` = everyone)."""
    if ranks is None:
        return True
    ranks = list(ranks)
    if -1 in ranks:
        return True
    return _current_rank() in ranks


def log_dist(message, ranks=None, level=logging.INFO):
    """Log ``message`` only on the listed ranks (reference: utils/logging.py log_dist)."""
    if should_log(ranks if ranks is not None else [0]):
        logger.log(level, message, stacklevel=2)


def print_json_dist(message: dict, ranks=None, path=None):
    """Dump a dict as json on the listed ranks (used by the autotuner metric hand-off)."""
    if should_log(ranks if ranks is not None else [0]):
        message = dict(message)
        message["rank"] = _current_rank()
        if path is None:
            print(json.dumps(message), flush=True)
        else:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            with open(path, "w") as f:
                json.dump(message, f)
                f.flush()


@functools.lru_cache(None)
def warning_once(msg: str):
    logger.warning(msg, stacklevel=2)


def get_log_level_from_string(s: str) -> int:
    if s.lower() not in _LEVELS:
        raise ValueError(f"unknown log level {s!r}; choose from {sorted(_LEVELS)}")
    return _LEVELS[s.lower()]


def set_log_level(level):
    if isinstance(level, str):
        level = get_log_level_from_string(level)
    logger.setLevel(level)


log_levels = {"debug": logging.DEBUG, "info": logging.INFO, "warning": logging.WARNING, "error": logging.ERROR,
              "critical": logging.CRITICAL}


def get_current_level():
    return logger.getEffectiveLevel()


def should_log_le(max_log_level_str):
    """Is the current level at most ``max_log_level_str`` (i.e. would a message of that level be shown)?"""
    if not isinstance(max_log_level_str, str):
        raise ValueError(f"{max_log_level_str} is not a string")
    key = max_log_level_str.lower()
    if key not in log_levels:
        raise ValueError(f"{max_log_level_str} is not one of the `logging` levels")
    return get_current_level() <= log_levels[key]


def print_configuration(args, name):
    logger.info(f"{name}:")
    for arg in sorted(vars(args)):
        logger.info(f"  {arg} {'.' * max(1, 29 - len(arg))} {getattr(args, arg)}")
