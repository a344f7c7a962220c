"""``comms_logger`` config block (reference: ``comm/config.py:11``)."""
from typing import List

from pydantic import BaseModel, ConfigDict


class CommsLoggerConfig(BaseModel):
    model_config = ConfigDict(extra="forbid")
    enabled: bool = False
    prof_all: bool = True
    prof_ops: List[str] = []
    verbose: bool = False
    debug: bool = False


class CommsConfig:

    def __init__(self, ds_config: dict):
        self.comms_logger = CommsLoggerConfig(**ds_config.get("comms_logger", {}))
        self.comms_logger_enabled = self.comms_logger.enabled


DeepSpeedCommsConfig = CommsConfig  # reference class name
