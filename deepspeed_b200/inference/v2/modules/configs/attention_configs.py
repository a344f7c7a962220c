"""Self-attention module config (reference ``modules/configs/attention_configs.py``)."""
from enum import Enum
from typing import Dict, Optional

from deepspeed_b200.inference.v2.inference_utils import DtypeEnum

from ..ds_module import DSModuleConfig


class PositionalEmbeddingType(Enum):
    none = "none"            # positions handled by the embedding layer
    rotate_half = "rotate_half"  # GPT-NeoX / llama rotary
    alibi = "alibi"


class RotateHalfConfig(DSModuleConfig):
    use_trained_freqs: bool = False
    theta_base: float = 10_000.0
    rotate_dim: Optional[int] = None  # None: the full head


class MaskingType(Enum):
    causal = "causal"
    local = "local"
    asymmetric = "asymmetric"  # caller-supplied mask


class DSSelfAttentionConfig(DSModuleConfig):
    n_heads_q: int
    n_heads_kv: int
    head_size: int
    max_sequences: int = 128
    scale_factor: float = 1.0
    input_dtype: DtypeEnum = DtypeEnum.fp16
    output_dtype: DtypeEnum = DtypeEnum.fp16
    masking_type: MaskingType = MaskingType.causal
    masking_args: Dict = {}
    positional_embedding_type: PositionalEmbeddingType = PositionalEmbeddingType.none
    positional_embedding_config: Optional[RotateHalfConfig] = None
