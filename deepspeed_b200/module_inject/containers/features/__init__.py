from .meta_tensor import MetaTensorContainer  # noqa: F401
from .megatron import MegatronContainer  # noqa: F401
from .split_qkv import HybridSplitQKVContainer  # noqa: F401
from .gated_mlp import HybridGatedMLPContainer  # noqa: F401
from .hybrid_engine import HybridEngineContainer  # noqa: F401
from .hybrid_megatron import HybridMegatronContainer  # noqa: F401
