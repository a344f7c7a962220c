"""``FP16_Optimizer`` (reference ``runtime/fp16/fused_optimizer.py:33``): fp16 parameters, flat fp32 master copy,
static or dynamic loss scaling, overflow-skipping step.  Unified implementation: ``ZeroShardedOptimizer`` stage 0
with ``model_dtype=fp16`` — the master copy *is* one flat arena stepped by the fused Adam/LAMB kernel."""
import torch

from deepspeed_b200.runtime.zero.sharded import ZeroShardedOptimizer


class FP16_Optimizer(ZeroShardedOptimizer):
    """Reference constructor signature over the unified optimizer (a class, so ``isinstance`` checks work)."""

    def __init__(self, init_optimizer, deepspeed=None, static_loss_scale=1.0, dynamic_loss_scale=False,
                 initial_dynamic_scale=2**32, dynamic_loss_args=None, verbose=True, mpu=None, clip_grad=0.0,
                 fused_adam_legacy=False, has_moe_layers=False, timers=None, module=None, dp_process_group=None,
                 gradient_accumulation_steps=1):
        module = module if module is not None else getattr(deepspeed, "module", None)
        assert module is not None, "pass module= or deepspeed= (engine)"
        args = dict(dynamic_loss_args or {})
        if dynamic_loss_scale:
            args.setdefault("init_scale", initial_dynamic_scale)
        super().__init__(module, 0, client_optimizer=init_optimizer, dp_group=dp_process_group, model_dtype=torch.float16,
                         gradient_accumulation_steps=gradient_accumulation_steps, gradient_clipping=clip_grad, mpu=mpu,
                         timers=timers, loss_scale_config={"dynamic": dynamic_loss_scale, "static_loss_scale": static_loss_scale,
                                                           "dynamic_args": args or None})

    @property
    def fp32_groups_flat(self):
        """fp32 master copy (reference: one flat tensor per param group; here one arena)."""
        return self.fp32_partitioned_groups_flat
