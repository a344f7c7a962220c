from . import constants as C


def get_aio_config(param_dict):
    d = (param_dict or {}).get(C.AIO) or {}
    return {
        C.AIO_BLOCK_SIZE: d.get(C.AIO_BLOCK_SIZE, C.AIO_BLOCK_SIZE_DEFAULT),
        C.AIO_QUEUE_DEPTH: d.get(C.AIO_QUEUE_DEPTH, C.AIO_QUEUE_DEPTH_DEFAULT),
        C.AIO_INTRA_OP_PARALLELISM: d.get(C.AIO_INTRA_OP_PARALLELISM, C.AIO_INTRA_OP_PARALLELISM_DEFAULT),
        C.AIO_SINGLE_SUBMIT: d.get(C.AIO_SINGLE_SUBMIT, C.AIO_SINGLE_SUBMIT_DEFAULT),
        C.AIO_OVERLAP_EVENTS: d.get(C.AIO_OVERLAP_EVENTS, C.AIO_OVERLAP_EVENTS_DEFAULT),
        C.AIO_USE_GDS: d.get(C.AIO_USE_GDS, C.AIO_USE_GDS_DEFAULT),
    }


def make_handle(aio_config, use_gds=None):
    """aio_handle (or gds_handle) from a config dict / model."""
    g = (lambda k, dflt: getattr(aio_config, k, dflt)) if not isinstance(aio_config, dict) else \
        (lambda k, dflt: aio_config.get(k, dflt))
    args = (g(C.AIO_BLOCK_SIZE, C.AIO_BLOCK_SIZE_DEFAULT), g(C.AIO_QUEUE_DEPTH, C.AIO_QUEUE_DEPTH_DEFAULT),
            g(C.AIO_SINGLE_SUBMIT, C.AIO_SINGLE_SUBMIT_DEFAULT), g(C.AIO_OVERLAP_EVENTS, C.AIO_OVERLAP_EVENTS_DEFAULT),
            g(C.AIO_INTRA_OP_PARALLELISM, C.AIO_INTRA_OP_PARALLELISM_DEFAULT))
    if use_gds if use_gds is not None else g(C.AIO_USE_GDS, False):
        from deepspeed_b200.ops.gds import gds_handle
        return gds_handle(*args)
    from deepspeed_b200.ops.aio import aio_handle
    return aio_handle(*args)
