"""``FP16_UnfusedOptimizer`` (reference ``runtime/fp16/unfused_optimizer.py:24``): per-tensor fp32 master weights
for optimizers that need per-tensor statistics (LAMB trust ratios).  Here: stage-0 sharded optimizer whose flat
optimizer runs in per-tensor mode (``FlatLamb.per_tensor``)."""
from .fused_optimizer import FP16_Optimizer


class FP16_UnfusedOptimizer(FP16_Optimizer):

    def __init__(self, init_optimizer, deepspeed=None, static_loss_scale=1.0, dynamic_loss_scale=False, dynamic_loss_args=None,
                 verbose=True, mpu=None, clip_grad=0.0, fused_lamb_legacy=False, **kw):
        super().__init__(init_optimizer, deepspeed=deepspeed, static_loss_scale=static_loss_scale,
                         dynamic_loss_scale=dynamic_loss_scale, dynamic_loss_args=dynamic_loss_args, mpu=mpu,
                         clip_grad=clip_grad, **kw)

    @property
    def fp32_groups(self):
        """Per-group lists of fp32 master tensors (reference attribute); the single arena here."""
        return [self.fp32_partitioned_groups_flat]
