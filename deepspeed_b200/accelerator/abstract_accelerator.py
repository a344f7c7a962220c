"""Reference module path of the accelerator ABC (``accelerator/abstract_accelerator.py``)."""
from . import DeepSpeedAccelerator  # noqa: F401
