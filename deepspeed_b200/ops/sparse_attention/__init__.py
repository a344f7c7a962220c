from .sparsity_config import (SparsityConfig, DenseSparsityConfig, FixedSparsityConfig, VariableSparsityConfig,  # noqa: F401
                              BigBirdSparsityConfig, BSLongformerSparsityConfig, LocalSlidingWindowSparsityConfig)
from .sparse_self_attention import SparseSelfAttention, block_sparse_attention  # noqa: F401
from .bert_sparse_self_attention import BertSparseSelfAttention  # noqa: F401
from .sparse_attention_utils import SparseAttentionUtils  # noqa: F401
from .matmul import MatMul  # noqa: F401
from .softmax import Softmax  # noqa: F401
