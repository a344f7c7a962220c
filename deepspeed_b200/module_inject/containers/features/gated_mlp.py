"""Mixin for gated MLPs (gate / up kept as two modules) in the hybrid engine (reference
``containers/features/gated_mlp.py``)."""
import torch


class HybridGatedMLPContainer:

    def set_mlp_gate(self):
        """Subclasses assign self.inter_up_w, self.inter_up_b, self.inter_gate_w, self.inter_gate_b."""
        raise NotImplementedError

    def refresh_fused_mlp(self):
        self.set_mlp_gate()
        with torch.no_grad():
            self.module.inter_w.copy_(torch.cat([self.inter_gate_w, self.inter_up_w], 0).to(self.module.inter_w.dtype))
