from .attention_configs import (DSSelfAttentionConfig, MaskingType, PositionalEmbeddingType, RotateHalfConfig)  # noqa: F401
from .embedding_config import DSEmbeddingsConfig  # noqa: F401
from .linear_config import DSLinearConfig  # noqa: F401
from .moe_config import DSMoEConfig  # noqa: F401
from .norm_config import DSNormConfig  # noqa: F401
from .unembed_config import DSUnembedConfig  # noqa: F401
