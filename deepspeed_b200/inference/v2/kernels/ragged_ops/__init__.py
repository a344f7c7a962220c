"""Ragged-batch kernels (reference ``inference/v2/kernels/ragged_ops``)."""
from .atom_builder.atom_builder import AtomBuilder  # noqa: F401
from .blocked_flash.blocked_flash import BlockedFlashAttn  # noqa: F401
from .embed.embed import RaggedEmbeddingKernel  # noqa: F401
from .linear_blocked_kv_rotary.blocked_kv_rotary import BlockedRotaryEmbeddings  # noqa: F401
from .linear_blocked_kv_rotary.blocked_trained_kv_rotary import BlockedTrainedRotaryEmbeddings  # noqa: F401
from .linear_blocked_kv_rotary.linear_blocked_kv_copy import LinearBlockedKVCopy  # noqa: F401
from .logits_gather.logits_gather import RaggedLogitsGather  # noqa: F401
from .moe_gather.moe_gather import MoEGather  # noqa: F401
from .moe_scatter.moe_scatter import MoEScatter  # noqa: F401
from .top_k_gating.top_k_gating import RaggedTopKGating  # noqa: F401
