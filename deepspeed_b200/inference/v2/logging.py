"""Logger of the inference engine (reference ``inference/v2/logging.py``)."""
import functools
import logging

from deepspeed_b200.utils.logging import LoggerFactory


@functools.lru_cache(maxsize=None)
def _make(level):
    log = LoggerFactory.create_logger(name="DS-Inference", level=level)
    log.debug("Inference logger created.")
    return log


def inference_logger(level: int = logging.INFO) -> logging.Logger:
    """The process-wide ``DS-Inference`` logger; the level given on first use sticks."""
    if _make.cache_info().currsize:
        return next(iter(_loggers()))
    return _make(level)


def _loggers():
    # lru_cache keeps exactly one entry (the first level asked for)
    yield logging.getLogger("DS-Inference")
