"""Package-level constants (reference ``deepspeed/constants.py``)."""
import os
from datetime import timedelta

TORCH_DISTRIBUTED_DEFAULT_PORT = 29500
default_pg_timeout = timedelta(minutes=int(os.getenv("DEEPSPEED_TIMEOUT", default=30)))
INFERENCE_GENERIC_MODE = "generic"
INFERENCE_SPECIALIZED_MODE = "specialized"
