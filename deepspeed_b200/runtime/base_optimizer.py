"""Marker bases for framework-owned optimizers (reference ``runtime/base_optimizer.py``)."""
import os

from deepspeed_b200.utils import logger


class DeepSpeedOptimizer:
    pass


class ZeROOptimizer(DeepSpeedOptimizer):
    """Base of every sharded optimizer; carries the universal-checkpoint loader entry point."""

    def load_hp_checkpoint_state_from_checkpoint_dir(self, lp_groups_name: str = None, checkpoint_dir: str = None) -> None:
        """Load per-parameter fp32 weights + optimizer moments from a *universal* checkpoint folder
        (``<dir>/zero/<param name>/{fp32,exp_avg,exp_avg_sq}.pt``)."""
        from deepspeed_b200.checkpoint.universal_checkpoint import load_universal_into_optimizer
        zero_dir = os.path.join(checkpoint_dir, "zero") if os.path.isdir(os.path.join(checkpoint_dir, "zero")) else checkpoint_dir
        logger.info(f"loading universal checkpoint state from {zero_dir}")
        load_universal_into_optimizer(self, zero_dir)
