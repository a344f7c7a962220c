from .attn import get_local_heads  # noqa: F401
from .attn_out import attn_out_in_features, shard_attn_out_param  # noqa: F401
from .embedding import shard_embedding_param, sharded_embedding_dim  # noqa: F401
from .mlp import shard_mlp_1_param, shard_mlp_2_param, sharded_intermediate_dim  # noqa: F401
from .qkv import qkv_out_features, shard_qkv_param  # noqa: F401
from .types import DEFAULT_SHARD_GRANULARITY, ShardingType  # noqa: F401
from .unembed import shard_unembed_param, sharded_unembed_dim  # noqa: F401
from .utils import get_shard_endpoints, shard_param  # noqa: F401
