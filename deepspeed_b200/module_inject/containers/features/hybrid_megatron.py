"""Hybrid-engine support for Megatron-layout layers (reference ``containers/features/hybrid_megatron.py``): when the
training layer stores its fused QKV per head (``[heads, 3, d]``), the refresh before a generation phase must de-interleave
again, because the optimizer keeps updating the per-head layout."""
from ...policy import deinterleave_qkv
from .hybrid_engine import HybridEngineContainer
from .megatron import MegatronContainer


class HybridMegatronContainer(MegatronContainer, HybridEngineContainer):

    def _align_qkv(self, x):
        return deinterleave_qkv(x, self.num_attention_heads) if getattr(self.policy, "is_megatron_v2", False) else x

    def refresh(self):
        # MegatronContainer.initialize_tensors de-interleaves on the way in; then slice + copy as usual
        super().refresh()
