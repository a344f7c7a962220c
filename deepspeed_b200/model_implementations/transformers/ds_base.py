import torch.nn as nn


class DeepSpeedTransformerBase(nn.Module):
    """Common ancestor of the per-family fused inference layers (reference ``ds_base.py`` placeholder)."""
