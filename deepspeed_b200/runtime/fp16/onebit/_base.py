import torch

from deepspeed_b200 import comm as dist


class _CompressedOptimizer(torch.optim.Optimizer):
    """Shared plumbing: the 1-bit collective, error-feedback buffers, and the engine switch that turns the
    dense gradient all-reduce off while the optimizer communicates compressed momentum itself."""

    def _setup(self, deepspeed, cuda_aware=False, comm_backend_name="nccl"):
        from deepspeed_b200.runtime.comm import CompressedBackend
        self.deepspeed = deepspeed
        self.comm_backend_name = comm_backend_name
        group = getattr(deepspeed, "seq_data_parallel_group", None) if deepspeed is not None else None
        if dist.is_initialized():
            self.comm_backend_handle = CompressedBackend(group=group)
            self.size = self.comm_backend_handle.size
        else:
            self.comm_backend_handle, self.size = None, 1
        self.divider = 8 * self.size
        self.using_pipeline = bool(deepspeed is not None and hasattr(deepspeed, "pipeline_enable_backward_allreduce"))

    def _set_engine_allreduce(self, on: bool):
        if self.deepspeed is None:
            return
        if self.using_pipeline:
            self.deepspeed.pipeline_enable_backward_allreduce = on
        self.deepspeed.enable_backward_allreduce = on

    def _error_buffers(self, state, p):
        if "worker_error" not in state:
            n = p.numel()
            padded = n if n % self.divider == 0 else n + self.divider - n % self.divider
            state["worker_error"] = torch.zeros(padded, dtype=torch.float32, device=p.device)
            state["server_error"] = torch.zeros(padded // self.size, dtype=torch.float32, device=p.device)
        return state["worker_error"], state["server_error"]

    def _compressed_mean(self, t, state, p):
        if self.size == 1 or self.comm_backend_handle is None:
            return t
        we, se = self._error_buffers(state, p)
        return self.comm_backend_handle.compressed_allreduce(t.float(), we, se).to(t.dtype)
