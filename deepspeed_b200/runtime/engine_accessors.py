"""Read-only accessors of ``DeepSpeedEngine`` over its config (reference ``runtime/engine.py:560-1010``: ``pld_enabled()``,
``zero_overlap_comm()``, ``autotuning_metric_path()`` ...).  They are generated from a table ``method name → attribute path
inside the engine config`` so the list is auditable in one place; the few that carry logic are written out below."""
import os

import torch

_DIRECT = (  # engine.<name>() == engine._config.<name>
    "checkpoint_tag_validation_enabled checkpoint_tag_validation_fail elasticity_enabled pld_enabled pld_params eigenvalue_enabled "
    "eigenvalue_verbose eigenvalue_max_iter eigenvalue_tol eigenvalue_stability eigenvalue_gas_boundary_resolution "
    "eigenvalue_layer_name eigenvalue_layer_num curriculum_enabled_legacy curriculum_params_legacy data_efficiency_enabled "
    "data_efficiency_config sparse_gradients_enabled optimizer_params optimizer_legacy_fusion scheduler_params "
    "zero_allow_untested_optimizer zero_force_ds_cpu_optimizer mics_shard_size graph_harvesting "
    "fp16_master_weights_and_gradients amp_params fp16_auto_cast use_node_local_storage load_universal_checkpoint "
    "gradient_predivide_factor initial_dynamic_scale dynamic_loss_scale_args aio_config dataloader_drop_last").split()

_NESTED = {  # engine.<name>() == engine._config.<section>.<field>
    "flops_profiler_config": {"flops_profiler_recompute_fwd_factor": "recompute_fwd_factor", "flops_profiler_module_depth": "module_depth",
                              "flops_profiler_top_modules": "top_modules", "flops_profiler_output_file": "output_file"},
    "zero_config": {"zero_reduce_scatter": "reduce_scatter", "zero_overlap_comm": "overlap_comm",
                    "zero_offload_optimizer": "offload_optimizer", "zero_offload_param": "offload_param",
                    "zero_sub_group_size": "sub_group_size", "zero_reduce_bucket_size": "reduce_bucket_size",
                    "zero_multi_rank_bucket_allreduce": "use_multi_rank_bucket_allreduce",
                    "zero_allgather_bucket_size": "allgather_bucket_size", "zero_contiguous_gradients": "contiguous_gradients",
                    "zero_load_from_fp32_weights": "load_from_fp32_weights", "zero_elastic_checkpoint": "elastic_checkpoint",
                    "zero_max_live_parameters": "max_live_parameters", "zero_max_reuse_distance": "max_reuse_distance",
                    "zero_prefetch_bucket_size": "prefetch_bucket_size",
                    "zero_module_granularity_threshold": "module_granularity_threshold",
                    "zero_param_persistence_threshold": "param_persistence_threshold",
                    "zero_model_persistence_threshold": "model_persistence_threshold",
                    "zero_gather_16bit_weights_on_model_save": "gather_16bit_weights_on_model_save",
                    "zero_grad_hooks": "grad_hooks", "zero_legacy_stage1": "legacy_stage1",
                    "zero_ignore_unused_parameters": "ignore_unused_parameters", "zero_allgather_partitions": "allgather_partitions",
                    "zero_round_robin_gradients": "round_robin_gradients", "zero_hpz_partition_size": "zero_hpz_partition_size",
                    "zero_quantized_weights": "zero_quantized_weights",
                    "zero_quantized_nontrainable_weights": "zero_quantized_nontrainable_weights",
                    "zero_quantized_gradients": "zero_quantized_gradients", "zeropp_loco_param": "zeropp_loco_param",
                    "zero_log_trace_cache_warnings": "log_trace_cache_warnings"},
    "tensor_parallel_config": {"autotp_size": "autotp_size"},
}
_NESTED_DEFAULTS = {"zero_grad_hooks": True, "zero_legacy_stage1": False}


def _field(obj, name, default=None):
    if obj is None:
        return default
    return obj.get(name, default) if isinstance(obj, dict) else getattr(obj, name, default)


class EngineConfigAccessors:
    """Mixin: every method reads ``self._config``; nothing here mutates engine state (except the
    ``communication_data_type`` setter, as in the reference)."""

    # ---- autotuning (the section is a plain dict here)
    def _cfg_at(self, key, default=None):
        return _field(getattr(self._config, "autotuning_config", None), key, default)

    def autotuning_enabled(self):
        return bool(self._cfg_at("enabled", False))

    def autotuning_start_profile_step(self):
        return self._cfg_at("start_profile_step", 3)

    def autotuning_end_profile_step(self):
        return self._cfg_at("end_profile_step", 5)

    def autotuning_metric_path(self):
        return self._cfg_at("metric_path") or os.path.join(os.getcwd(), "autotuning_metric.json")

    def autotuning_model_info_path(self):
        return self._cfg_at("model_info_path") or os.path.join(os.getcwd(), "autotuning_model_info.json")

    def autotuning_metric(self):
        return self._cfg_at("metric", "throughput")

    def autotuning_profile_model_info(self):
        info = self._cfg_at("model_info")
        return bool(self.autotuning_enabled() and info and info.get("profile", False))

    # ---- flops profiler (autotuning overrides)
    def flops_profiler_enabled(self):
        return bool(self._config.flops_profiler_config.enabled or self.autotuning_enabled())

    def flops_profiler_profile_step(self):
        return self.autotuning_start_profile_step() if self.autotuning_enabled() else self._config.flops_profiler_config.profile_step

    def flops_profiler_detailed(self):
        return False if self.autotuning_enabled() else self._config.flops_profiler_config.detailed

    # ---- data efficiency sections
    def _cfg_de(self, *path):
        cur = self._config.data_efficiency_config or {}
        for k in path:
            cur = cur.get(k, {}) if isinstance(cur, dict) else {}
        return cur

    def data_sampling_enabled(self):
        return bool(self._cfg_de("data_sampling").get("enabled", False))

    def data_sampling_config(self):
        return self._cfg_de("data_sampling")

    def curriculum_learning_enabled(self):
        return bool(self._cfg_de("data_sampling", "curriculum_learning").get("enabled", False))

    def curriculum_learning_config(self):
        return self._cfg_de("data_sampling", "curriculum_learning")

    def random_ltd_enabled(self):
        return bool(self._cfg_de("data_routing", "random_ltd").get("enabled", False))

    def random_ltd_config(self):
        return self._cfg_de("data_routing", "random_ltd")

    # ---- PLD
    def pld_theta(self):
        return self.pld_params()["theta"]

    def pld_gamma(self):
        return self.pld_params()["gamma"]

    def get_pld_theta(self):
        pld = getattr(self, "progressive_layer_drop", None)
        return pld.get_theta() if pld else None

    # ---- ZeRO offload flavours
    def _offload_device(self):
        off = self._config.zero_config.offload_optimizer
        dev = getattr(off, "device", None) if off is not None else None
        return str(getattr(dev, "value", dev)) if dev is not None else "none"

    def zero_use_cpu_optimizer(self):
        return self._offload_device() in ("cpu", "nvme")

    def zero_cpu_offload(self):
        return self._offload_device() == "cpu"

    def zero_partial_offload(self):
        return getattr(self._config.zero_config.offload_optimizer, "ratio", 1.0)

    def zero_nvme_offload_optimizer(self):
        return self._offload_device() == "nvme"

    # ---- misc
    def is_elastic_model_parallel_supported(self):
        """Elastic training re-shapes only the data-parallel dimension unless elasticity v0.2 with a model-parallel size
        that divides the GPUs of a node is configured."""
        if not self.elasticity_enabled():
            return False
        params = getattr(self._config, "elasticity_params", None) or {}
        mp, per_node = params.get("model_parallel_size", 1), params.get("num_gpus_per_node", 1)
        return float(params.get("version", 0.1)) >= 0.2 and mp >= 1 and per_node % mp == 0

    def quantize_training(self):
        """MoQ knobs, in the order the reference returns them."""
        wq = (getattr(self._config, "compression_config", None) or {}).get("weight_quantization", {})
        sh = wq.get("shared_parameters", {})
        mixed = sh.get("fp16_mixed_quantize", {}) or {}
        first = next(iter((wq.get("different_groups") or {}).values()), {}).get("params", {})
        return (sh.get("quantize_weight_in_forward", False), sh.get("enabled", False), sh.get("quantize_groups", 1),
                mixed.get("enabled", False), mixed.get("quantize_change_ratio", 0.001), sh.get("quantization_type", "symmetric"),
                sh.get("rounding", "nearest"), sh.get("quantize_verbose", False), sh.get("quantizer_kernel", False),
                first.get("start_bits"), first.get("target_bits"), first.get("quantization_period", 1))

    def swap_tensor_config(self):
        return getattr(self._config, "swap_tensor_config", None) or getattr(self._config.zero_config, "offload_optimizer", None)

    def postscale_gradients(self):
        return not self._config.prescale_gradients

    @property
    def communication_data_type(self):
        res = self._config.communication_data_type
        if res is not None:
            return res
        if self.fp16_enabled():
            return torch.float16
        if self.bfloat16_enabled():
            return torch.bfloat16
        return torch.float32

    @communication_data_type.setter
    def communication_data_type(self, value):
        self._config.communication_data_type = value

    @staticmethod
    def is_map_style_dataset(obj):
        return hasattr(obj, "__getitem__") and hasattr(obj, "__len__")

    @staticmethod
    def is_iterable_style_dataset(obj):
        return isinstance(obj, torch.utils.data.IterableDataset)

    # ---- sparse-gradient collectives (embedding gradients as (indices, values))
    def all_gather_scalar(self, value, dp_group):
        from deepspeed_b200 import comm as dist
        out = [value.new_zeros(value.size()) for _ in range(dist.get_world_size(group=dp_group))]
        dist.all_gather(out, value, group=dp_group)
        return out

    def sparse_all_gather(self, value, dp_group):
        """All-gather of per-rank different-length tensors: pad to the longest, gather, trim."""
        from deepspeed_b200 import comm as dist
        n = torch.tensor([value.size(0)], dtype=torch.long, device=value.device)
        sizes = [int(s) for s in torch.cat(self.all_gather_scalar(n, dp_group)).tolist()]
        longest = max(sizes)
        pad = value.new_zeros((longest - value.size(0), ) + tuple(value.shape[1:]))
        padded = torch.cat([value, pad])
        out = [torch.empty_like(padded) for _ in sizes]
        dist.all_gather(out, padded, group=dp_group)
        return [t[:s] for t, s in zip(out, sizes)]

    def sparse_allreduce(self, sparse, dp_group, dp_world_size=None):
        from deepspeed_b200 import comm as dist
        w = dp_world_size or dist.get_world_size(group=dp_group)
        vals = sparse.values.float() if sparse.values.dtype != torch.float32 else sparse.values
        if self.postscale_gradients():
            if self.gradient_predivide_factor() != 1.0:
                vals = vals / self.gradient_predivide_factor()
                post = self.gradient_predivide_factor() / w
            else:
                post = 1.0 / w
        else:
            vals, post = vals / w, 1.0
        sparse.indices = torch.cat(self.sparse_all_gather(sparse.indices, dp_group))
        sparse.values = torch.cat(self.sparse_all_gather(vals, dp_group)) * post
        return sparse

    def sparse_allreduce_bucket(self, bucket, dp_group, dp_world_size=None):
        return [self.sparse_allreduce(s, dp_group, dp_world_size) for s in bucket]

    def sparse_allreduce_no_retain(self, bucket, dp_group, dp_world_size=None):
        """Reduce sparse gradients and write the densified result back into the owning tensors."""
        for src, red in zip(bucket, self.sparse_allreduce_bucket(bucket, dp_group, dp_world_size)):
            target = getattr(src, "orig_dense_tensor", None)
            if target is not None:
                target.data.copy_(red.to_dense().to(target.dtype))

    def allreduce_no_retain(self, bucket, dp_group, numel_per_bucket=500000000, dp_world_size=None):
        """Bucketed dense all-reduce (average) of a list of tensors, in place."""
        from deepspeed_b200 import comm as dist
        w = dp_world_size or dist.get_world_size(group=dp_group)
        small, numel = [], 0

        def flush():
            if not small:
                return
            flat = torch.cat([t.reshape(-1) for t in small])
            if self.postscale_gradients():
                dist.all_reduce(flat, group=dp_group)
                flat.div_(w)
            else:
                flat.div_(w)
                dist.all_reduce(flat, group=dp_group)
            off = 0
            for t in small:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()
            small.clear()

        for t in bucket:
            small.append(t)
            numel += t.numel()
            if numel > numel_per_bucket:
                flush()
                numel = 0
        flush()

    def print_forward_breakdown(self, fwd_time):
        """MoE forward breakdown (gate / experts / all-to-all) from the layers' own timers."""
        gate = moe = a2a = 0.0
        for m in self.module.modules():
            if hasattr(m, "gate_time") or hasattr(m, "time_moe"):
                gate += getattr(m, "gate_time", 0.0)
                moe += getattr(m, "time_moe", 0.0)
                a2a += getattr(m, "time_falltoall", 0.0) + getattr(m, "time_salltoall", 0.0)
        from deepspeed_b200.utils.logging import log_dist
        log_dist(f"time (ms) | fwd: {fwd_time:.2f} (fwd_moe: {moe:.2f}, 1st_a2a+2nd_a2a: {a2a:.2f}, top_k: {gate:.2f})", ranks=[0])


def _install():
    def direct(name):
        return lambda self: getattr(self._config, name)

    def nested(section, field, name):
        return lambda self: _field(getattr(self._config, section), field, _NESTED_DEFAULTS.get(name))

    for n in _DIRECT:
        if n not in EngineConfigAccessors.__dict__:
            f = direct(n)
            f.__name__ = n
            setattr(EngineConfigAccessors, n, f)
    for section, fields in _NESTED.items():
        for n, fld in fields.items():
            f = nested(section, fld, n)
            f.__name__ = n
            setattr(EngineConfigAccessors, n, f)


_install()
