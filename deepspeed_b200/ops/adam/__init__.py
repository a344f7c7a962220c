from .fused_adam import FusedAdam  # noqa: F401
from .cpu_adam import DeepSpeedCPUAdam  # noqa: F401
