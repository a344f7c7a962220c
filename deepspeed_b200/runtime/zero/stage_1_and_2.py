"""ZeRO stage 1 / 2 optimizer under the reference's class name and constructor
(reference ``runtime/zero/stage_1_and_2.py:DeepSpeedZeroOptimizer``).

There is one sharded optimizer in this framework (``sharded.ZeroShardedOptimizer``, stage is a parameter); this class
adapts the reference's module-less signature - it receives only the client optimizer's param groups - by hanging the
parameters on a container module so the unit planner can walk them.
"""
import torch
from torch import nn

from deepspeed_b200.runtime.zero.config import DeepSpeedZeroConfig
from deepspeed_b200.runtime.zero.mem_estimator import (  # noqa: F401  (the reference defines these here)
    estimate_zero2_model_states_mem_needs, estimate_zero2_model_states_mem_needs_all_cold,
    estimate_zero2_model_states_mem_needs_all_live)
from deepspeed_b200.runtime.zero._stage_helpers import (  # noqa: F401,A004
    input, split_half_float_double, isclose, lcm, get_alignment_padding, print_rank_msg, model_to_params, move_to_cpu, pg_correctness_test, OPTIMIZER_TIMERS)
from deepspeed_b200.runtime.zero.sharded import ZeroShardedOptimizer


class _ParamBag(nn.Module):
    """Container giving module-less parameter lists a module identity (one planner unit per param group)."""

    def __init__(self, param_groups, param_names=None):
        super().__init__()
        names = param_names or {}
        for gi, g in enumerate(param_groups):
            holder = nn.Module()
            for pi, p in enumerate(g["params"]):
                holder.register_parameter(str(names.get(p, f"p{pi}")).replace(".", "_"), p)
            self.add_module(f"group{gi}", holder)


def _dtype_of(param_groups):
    for g in param_groups:
        for p in g["params"]:
            return p.dtype
    return torch.float32


class DeepSpeedZeroOptimizer(ZeroShardedOptimizer):

    def __init__(self, init_optimizer, param_names=None, timers=None, static_loss_scale=1.0, dynamic_loss_scale=False,
                 dynamic_loss_args=None, verbose=True, contiguous_gradients=True, reduce_bucket_size=500000000,
                 use_multi_rank_bucket_allreduce=True, allgather_bucket_size=5000000000, dp_process_group=None,
                 expert_parallel_group=None, expert_data_parallel_group=None, reduce_scatter=True, overlap_comm=False,
                 offload_optimizer_config=None, mpu=None, clip_grad=0.0, gradient_accumulation_dtype=torch.float32,
                 communication_data_type=torch.float16, postscale_gradients=True, gradient_predivide_factor=1.0,
                 gradient_accumulation_steps=1, ignore_unused_parameters=True, partition_grads=True, round_robin_gradients=False,
                 has_moe_layers=False, fp16_master_weights_and_gradients=False, elastic_checkpoint=False, module=None,
                 device=None):
        stage = 2 if partition_grads else 1
        zc = DeepSpeedZeroConfig(stage=stage, contiguous_gradients=contiguous_gradients, reduce_bucket_size=reduce_bucket_size,
                                 allgather_bucket_size=allgather_bucket_size, reduce_scatter=reduce_scatter,
                                 overlap_comm=overlap_comm, round_robin_gradients=round_robin_gradients,
                                 ignore_unused_parameters=ignore_unused_parameters,
                                 **({"offload_optimizer": offload_optimizer_config} if offload_optimizer_config else {}))
        mdt = _dtype_of(init_optimizer.param_groups)
        super().__init__(module if module is not None else _ParamBag(init_optimizer.param_groups, param_names), stage,
                         client_optimizer=init_optimizer, zero_config=zc, dp_group=dp_process_group, model_dtype=mdt,
                         grad_accum_dtype=gradient_accumulation_dtype if gradient_accumulation_dtype != torch.float32 else None,
                         gradient_accumulation_steps=gradient_accumulation_steps, gradient_clipping=clip_grad,
                         loss_scale_config={"static_loss_scale": static_loss_scale, "dynamic": dynamic_loss_scale,
                                            "dynamic_args": dynamic_loss_args},
                         communication_data_type=communication_data_type if communication_data_type != torch.float16 or
                         mdt == torch.float16 else None, prescale_gradients=not postscale_gradients,
                         gradient_predivide_factor=gradient_predivide_factor, mpu=mpu, timers=timers, device=device)
        self.partition_gradients = partition_grads
        self.has_moe_layers = has_moe_layers
        self.elastic_checkpoint = elastic_checkpoint
