"""Tiered ("nebula"-style) checkpointing config: fast local snapshot + background persistence with version retention."""
