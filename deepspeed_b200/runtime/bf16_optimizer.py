"""``BF16_Optimizer`` (reference ``runtime/bf16_optimizer.py:35``): bf16 parameters with fp32 master weights /
gradient accumulation and sharded optimizer state (ZeRO-1 semantics), also the optimizer used under pipeline
parallelism.  In this framework it is the unified ``ZeroShardedOptimizer`` at stage 0/1 with ``model_dtype=bf16``."""
import torch

from deepspeed_b200.runtime.zero.sharded import ZeroShardedOptimizer


class BF16_Optimizer(ZeroShardedOptimizer):
    """Reference constructor signature over the unified optimizer (a class, so ``isinstance`` checks work)."""

    def __init__(self, init_optimizer, param_names=None, mpu=None, clip_grad=0.0, norm_type=2, allgather_bucket_size=5e9,
                 dp_process_group=None, timers=None, grad_acc_dtype=None, graph_harvesting=False, immediate_grad_update=False,
                 has_moe_layers=False, module=None, gradient_accumulation_steps=1, stage=1):
        assert module is not None, "pass module= (the parameters' owner) so units can be planned"
        super().__init__(module, stage, client_optimizer=init_optimizer, dp_group=dp_process_group, model_dtype=torch.bfloat16,
                         grad_accum_dtype=grad_acc_dtype or torch.float32,
                         gradient_accumulation_steps=gradient_accumulation_steps, gradient_clipping=clip_grad, mpu=mpu,
                         timers=timers)


def print_rank_0(message, debug=False, force=False):
    from deepspeed_b200 import comm as dist
    if (debug or force) and (not dist.is_initialized() or dist.get_rank() == 0):
        print(message)
