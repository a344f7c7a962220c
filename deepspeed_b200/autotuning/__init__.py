from .autotuner import Autotuner  # noqa: F401
from .config import DeepSpeedAutotuningConfig  # noqa: F401
