from .blocked_allocator import BlockedAllocator  # noqa: F401
from .kv_cache import BlockedKVCache, split_kv  # noqa: F401
from .sequence_descriptor import DSSequenceDescriptor, PlaceholderSequenceDescriptor  # noqa: F401
from .ragged_wrapper import RaggedBatchWrapper  # noqa: F401
from .ragged_manager import DSStateManager  # noqa: F401
