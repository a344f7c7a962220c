"""Pluggable module layer of the ragged engine (reference ``inference/v2/modules``)."""
from . import heuristics  # noqa: F401
from .configs import *  # noqa: F401,F403
from .interfaces import *  # noqa: F401,F403
