from .attention_base import DSSelfAttentionBase, DSSelfAttentionRegistry  # noqa: F401
from .embedding_base import DSEmbeddingBase, DSEmbeddingRegistry  # noqa: F401
from .linear_base import DSLinearBase, DSLinearRegistry  # noqa: F401
from .moe_base import DSMoEBase, DSMoERegistry  # noqa: F401
from .post_norm_base import DSPostNormBase, DSPostNormRegistry  # noqa: F401
from .pre_norm_base import DSPreNormBase, DSPreNormRegistry  # noqa: F401
from .unembed_base import DSUnembedBase, DSUnembedRegistry  # noqa: F401
