"""AIO config keys (reference ``runtime/swap_tensor/constants.py``).

Names are declared section by section through ``_declare`` (later sections may refer to earlier names)."""


def _declare(**names):
    globals().update(names)
    return names


_declare(
    AIO="aio",
    AIO_BLOCK_SIZE="block_size",
    AIO_BLOCK_SIZE_DEFAULT=1048576,
    AIO_QUEUE_DEPTH="queue_depth",
    AIO_QUEUE_DEPTH_DEFAULT=8,
    AIO_INTRA_OP_PARALLELISM="intra_op_parallelism",
    AIO_INTRA_OP_PARALLELISM_DEFAULT=1,
    AIO_SINGLE_SUBMIT="single_submit",
    AIO_SINGLE_SUBMIT_DEFAULT=False,
    AIO_OVERLAP_EVENTS="overlap_events",
    AIO_OVERLAP_EVENTS_DEFAULT=True,
    AIO_USE_GDS="use_gds",
    AIO_USE_GDS_DEFAULT=False,
)
