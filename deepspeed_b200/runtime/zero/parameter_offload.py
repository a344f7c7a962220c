"""``DeepSpeedZeRoOffload``: ZeRO-3 parameter partitioning + fetch/release hooks *without* an optimizer (ZeRO-Inference,
or a frozen sub-model).  Reference: ``runtime/zero/parameter_offload.py:77`` (``setup_zero_stage3_hooks :232``,
``pre_sub_module_forward_function :442``, ``post_sub_module_forward_function :458``, ``mark_persistent_parameters :194``)
and the coordinator calls ``fetch_sub_module`` / ``release_sub_module`` (``partitioned_param_coordinator.py:276,412``).

Here the unit planner + hooks live in ``ZeroShardedOptimizer``; this class is that object configured with a no-op optimizer
and the reference's method names mapped onto unit operations.
"""
import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.runtime.zero.sharded import ZeroShardedOptimizer


class DeepSpeedZeRoOffload:

    def __init__(self, module, timers=None, ds_config=None, overlap_comm=True, prefetch_bucket_size=50_000_000,
                 max_reuse_distance=1_000_000_000, max_live_parameters=1_000_000_000, param_persistence_threshold=100_000,
                 model_persistence_threshold=None, dp_process_group=None, offload_param_config=None, mpu=None,
                 zero_param_parallel_group=None, zero_quantized_weights=False, zero_quantized_nontrainable_weights=False,
                 zero_module_granularity_threshold=0, log_trace_cache_warnings=False, dtype=None):
        zc = getattr(ds_config, "zero_config", None)
        if dtype is None:
            p = next(module.parameters())
            dtype = p.dtype
        self.module = module
        self.dtype = dtype
        self._zo = ZeroShardedOptimizer(module, 3, optimizer_name="sgd", optimizer_params={"lr": 0.0},
                                        param_groups=[{"params": []}], zero_config=zc, dp_group=dp_process_group,
                                        model_dtype=dtype, mpu=mpu, timers=timers)
        self.persistent_parameters = self.mark_persistent_parameters(param_persistence_threshold, model_persistence_threshold)
        self.forward_hooks = list(getattr(self._zo, "_hook_handles", []))
        self.backward_hooks = []

    # ---- reference method names --------------------------------------------------------------------------------
    def _rt_of(self, sub_module):
        for rt in self._zo.rts:
            if rt.u.module is sub_module:
                return rt
        for p in sub_module.parameters(recurse=False):
            rt = self._zo.unit_of_param.get(id(p))
            if rt is not None:
                return rt
        return None

    def mark_persistent_parameters(self, param_threshold, model_threshold=None):
        """Units small enough to stay gathered (unit-granular version of the reference's per-parameter rule)."""
        out = []
        for rt in self._zo.rts:
            if rt.u.persistent:
                out.extend(s.param for s in rt.u.slots)
        return out

    def setup_zero_stage3_hooks(self):
        return self.forward_hooks  # registered by the optimizer object at construction

    def fetch_sub_module(self, sub_module, forward=True):
        rt = self._rt_of(sub_module)
        if rt is not None:
            self._zo.fetch_unit(rt, forward=forward)

    def release_sub_module(self, sub_module, forward=True):
        rt = self._rt_of(sub_module)
        if rt is not None:
            self._zo.release_unit(rt)

    def pre_sub_module_forward_function(self, sub_module):
        self.fetch_sub_module(sub_module, forward=True)

    def post_sub_module_forward_function(self, sub_module):
        self.release_sub_module(sub_module, forward=True)

    def pre_sub_module_backward_function(self, sub_module):
        self.fetch_sub_module(sub_module, forward=False)

    def post_sub_module_backward_function(self, sub_module):
        self.release_sub_module(sub_module, forward=False)

    def partition_all_parameters(self):
        for rt in self._zo.rts:
            self._zo.release_unit(rt)

    def get_param_coordinator(self):
        return self._zo

    def empty_partition_cache(self):
        self.partition_all_parameters()

    def destroy(self):
        if hasattr(self._zo, "destroy"):
            self._zo.destroy()

    _remove_module_hooks = destroy


class ZeROOrderedDict(dict):
    """``module._parameters`` replacement that fetches a partitioned parameter on first touch (reference
    ``parameter_offload.py:45``): code that reaches into ``module._parameters[name]`` outside ``forward`` (weight tying,
    custom init) sees the full tensor instead of the empty placeholder."""

    def __init__(self, parent_module, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._parent_module = parent_module
        self._in_forward = False

    def __reduce__(self):
        return (dict, (), None, None, iter(self.items()))

    def __getitem__(self, key):
        param = super().__getitem__(key)
        if param is None or not hasattr(param, "ds_status"):
            return param
        from .partition_parameters import ZeroParamStatus
        if param.ds_status == ZeroParamStatus.NOT_AVAILABLE and not self._in_forward:
            if getattr(self._parent_module, "_parameters", None) is not None and hasattr(param, "all_gather"):
                param.all_gather()
        return param


def _inject_parameters(module, cls):
    """Swap every sub-module's ``_parameters`` dict for ``cls(parent_module=...)`` keeping the entries."""
    for m in module.modules():
        new = cls(parent_module=m)
        for k, v in m._parameters.items():
            new[k] = v
        m._parameters = new
