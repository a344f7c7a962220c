"""Reference-shaped facade over the gather / prefetch / release machinery of ``sharded.ZeroShardedOptimizer``
(reference ``runtime/zero/partitioned_param_coordinator.py``).

The reference coordinator records a per-submodule trace, prefetches ``__prefetch_bucket_sz`` parameters ahead and
releases by reuse distance.  Here those jobs belong to the optimizer itself at *unit* granularity (one flat all-gather
per layer, forward order traced on the first iteration, ``b200_prefetch_depth`` units kept in flight, a fixed pool of
gather buffers instead of max_live accounting); this class exposes that state through the reference's API so external
code (custom layers calling ``fetch_sub_module``, tools inspecting the trace) keeps working.
"""
from enum import Enum

from deepspeed_b200.runtime.zero.sharded import GATHERED, INFLIGHT, NOT_GATHERED


class ZeRoTraceMode(Enum):
    RECORD = 1  # first iteration: the forward unit order is being recorded
    COMPLETE = 2  # order known: prefetching follows it
    INVALID = 3  # the module graph changed; the order will be re-recorded


class InflightParamRegistry(dict):
    """param -> handle of its in-flight all-gather (the unit's CUDA event here)."""

    def __setitem__(self, param, handle):
        if param in self:
            raise RuntimeError(f"{param.ds_summary() if hasattr(param, 'ds_summary') else param} already in registry")
        super().__setitem__(param, handle)


class PartitionedParameterCoordinator:

    def __init__(self, zero_optimizer, prefetch_bucket_sz=None, max_reuse_distance_in_numel=None,
                 max_available_parameters_in_numel=None, allgather_stream=None, inflight_param_registry=None,
                 prefetch_nvme=False, timers=None, zero_quantized_weights=False, zero_quantized_nontrainable_weights=False,
                 fast_sharding_for_leaf_module=False, log_trace_cache_warnings=False):
        self.zo = zero_optimizer
        self._invalid = False
        self.__step_id = 0

    # ---- trace ---------------------------------------------------------------------------------------------------
    @property
    def trace_mode(self):
        if self._invalid:
            return ZeRoTraceMode.INVALID
        return ZeRoTraceMode.COMPLETE if self.zo._trace_done else ZeRoTraceMode.RECORD

    def is_complete_trace(self):
        return self.trace_mode is ZeRoTraceMode.COMPLETE

    def is_invalid_trace(self):
        return self.trace_mode is ZeRoTraceMode.INVALID

    def is_record_trace(self):
        return self.trace_mode is ZeRoTraceMode.RECORD

    def trace_prologue(self, sub_module):
        """Called before a submodule runs: detects a graph that diverged from the recorded order."""
        rt = self._unit_of(sub_module)
        if rt is None or not self.zo._trace_done:
            return
        order = self.zo._trace
        if self.__step_id < len(order) and order[self.__step_id] != rt.u.index:
            self._invalidate_trace()

    def record_module(self, sub_module):
        rt = self._unit_of(sub_module)
        if rt is not None and not self.zo._trace_done:
            self.zo._trace.append(rt.u.index)
        self.__step_id += 1

    def construct_parameter_trace_from_module_trace(self):
        """Parameter order implied by the recorded unit order."""
        return [s.param for i in self.zo._trace for s in self.zo.rts[i].u.slots]

    def _invalidate_trace(self):
        self._invalid = True
        self.zo.reset_step()

    def reset_step(self):
        """End of an iteration: a recorded trace becomes final; an invalidated one is recorded again."""
        self.__step_id = 0
        self._invalid = False

    # ---- fetch / release -------------------------------------------------------------------------------------------
    def _unit_of(self, sub_module):
        for rt in self.zo.rts:
            if rt.u.module is sub_module:
                return rt
        for p in sub_module.parameters(recurse=False):
            rt = self.zo.unit_of_param.get(id(p))
            if rt is not None:
                return rt
        return None

    def _units_under(self, module):
        seen, out = set(), []
        for p in module.parameters():
            rt = self.zo.unit_of_param.get(id(p))
            if rt is not None and id(rt) not in seen:
                seen.add(id(rt))
                out.append(rt)
        return out

    def fetch_sub_module(self, current_submodule, forward=True):
        """Block until the submodule's parameters are available on the compute stream; kicks off the prefetches."""
        rt = self._unit_of(current_submodule)
        if rt is not None:
            self.zo.fetch_unit(rt, forward=forward)
        self.__step_id += 1

    def release_sub_module(self, submodule, forward=False):
        rt = self._unit_of(submodule)
        if rt is not None:
            self.zo.release_unit(rt)

    def release_and_reset_all(self, module):
        """Drop every gathered / in-flight parameter under ``module`` (e.g. after an exception mid-forward)."""
        for rt in self._units_under(module):
            if rt.state == INFLIGHT and rt.gather_event is not None and self.zo.on_cuda:
                rt.gather_event.synchronize()
                rt.state = GATHERED
            self.zo.release_unit(rt)
        self.reset_step()

    # ---- introspection ---------------------------------------------------------------------------------------------
    @property
    def inflight_params(self):
        return [s.param for rt in self.zo.rts if rt.state == INFLIGHT for s in rt.u.slots]

    @property
    def available_parameter_numel(self):
        return sum(rt.u.full_numel for rt in self.zo.rts if rt.state != NOT_GATHERED and not rt.u.persistent)


def debug_rank0(message: str) -> None:
    from deepspeed_b200 import comm as dist
    from deepspeed_b200.utils.logging import logger
    if not dist.is_initialized() or dist.get_rank() == 0:
        logger.debug(message)


def get_all_parameters(sub_module, recurse=False):
    """Own + externally registered parameters of a module (reference ``partitioned_param_coordinator.py:31``)."""
    import itertools
    ext = getattr(sub_module, "_external_params", None) or getattr(sub_module, "ds_external_parameters", lambda: ())
    ext_items = ext.items() if isinstance(ext, dict) else (ext() if callable(ext) else ext)
    return itertools.chain(sub_module.named_parameters(recurse=recurse), ext_items)


def iter_params(module, recurse=False):
    return (p for _, p in get_all_parameters(module, recurse))
