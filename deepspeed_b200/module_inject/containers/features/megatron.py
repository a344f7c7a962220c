"""Mixin for Megatron checkpoints whose fused QKV is stored per head (``[heads, 3, d]``, "megatron v2")
(reference ``containers/features/megatron.py``)."""
from ...policy import deinterleave_qkv


class MegatronContainer:

    def initialize_tensors(self, enable_training=False):
        super().initialize_tensors(enable_training)
        if getattr(self.policy, "is_megatron_v2", False):
            self.qkvw = deinterleave_qkv(self.qkvw, self.num_attention_heads)
            self.qkvb = deinterleave_qkv(self.qkvb, self.num_attention_heads)
