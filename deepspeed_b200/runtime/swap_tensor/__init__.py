from .aio_config import get_aio_config  # noqa: F401
from .utils import SwapBuffer, SwapBufferPool, SwapBufferManager, swap_in_tensors, swap_out_tensors, MIN_AIO_BYTES  # noqa: F401
from .async_swapper import AsyncTensorSwapper  # noqa: F401
from .optimizer_utils import SwappedFlatState, FlatStateSwapper  # noqa: F401
from .partitioned_param_swapper import AsyncPartitionedParameterSwapper, PartitionedParamStatus  # noqa: F401
