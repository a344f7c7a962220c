"""Comms logger (reference ``utils/comms_logging.py``); implemented next to the comm facade in ``comm/comms_logging.py``."""
from deepspeed_b200.comm.comms_logging import *  # noqa: F401,F403
from deepspeed_b200.comm.comms_logging import CommsLogger  # noqa: F401
