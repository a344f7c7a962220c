"""Chunked multi-tensor launcher protocol (reference ``ops/adam/multi_tensor_apply.py``, from apex).

The fused optimizers here run over flat arenas, so they do not need it; it is kept for user kernels written against the
``op(chunk_size, noop_flag, tensor_lists, *args)`` convention."""


class MultiTensorApply:

    def __init__(self, chunk_size):
        self.chunk_size = chunk_size

    def __call__(self, op, noop_flag_buffer, tensor_lists, *args):
        return op(self.chunk_size, noop_flag_buffer, tensor_lists, *args)
