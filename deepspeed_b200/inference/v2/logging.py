"""Logger of the inference engine (reference ``inference/v2/logging.py``)."""
import logging

from deepspeed_b200.utils.logging import LoggerFactory

inf_logger = None


def inference_logger(level: int = logging.INFO) -> logging.Logger:
    """The process-wide ``DS-Inference`` logger; the level given on first use sticks."""
    global inf_logger
    if inf_logger is None:
        inf_logger = LoggerFactory.create_logger(name="DS-Inference", level=level)
        inf_logger.debug("Inference logger created.")
    return inf_logger
