"""Mixture-of-Experts with expert parallelism (reference ``deepspeed/moe``)."""
from .layer import MoE  # noqa: F401
