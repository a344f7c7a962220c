from .compress import init_compression, redundancy_clean, student_initialization  # noqa: F401
from .scheduler import compression_scheduler  # noqa: F401
from .basic_layer import (LinearLayer_Compress, Conv2dLayer_Compress, Embedding_Compress, BNLayer_Compress,  # noqa: F401
                          QuantAct, ColumnParallelLinear_Compress, RowParallelLinear_Compress)
from .helper import convert_conv1d_to_linear  # noqa: F401
