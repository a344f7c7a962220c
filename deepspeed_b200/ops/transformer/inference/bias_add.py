"""``nhwc_bias_add`` (reference ``ops/transformer/inference/bias_add.py``): channels-last bias add for diffusers conv
blocks, optionally fused with a second (biased) tensor."""
from typing import Optional

import torch

from deepspeed_b200.ops.spatial import nhwc_bias_add as _nhwc_bias_add  # handles logical-NCHW channels_last tensors


def nhwc_bias_add(activation: torch.Tensor, bias: torch.Tensor, other: Optional[torch.Tensor] = None,
                  other_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    return _nhwc_bias_add(activation, bias, other, other_bias)
