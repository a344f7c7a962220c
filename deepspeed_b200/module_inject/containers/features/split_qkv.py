"""Mixin for families that keep q/k/v as separate modules: lets the hybrid engine alias the fused QKV rows back onto
the three original parameters (reference ``containers/features/split_qkv.py``)."""
import torch


class HybridSplitQKVContainer:

    def set_q_k_v(self):
        """Subclasses assign self.qw, self.qb, self.kw, self.kb, self.vw, self.vb from the original layer."""
        raise NotImplementedError

    def attention_qkv_views(self):
        """Row ranges of the fused QKV weight belonging to q, k and v."""
        self.set_q_k_v()
        nq, nk = self.qw.shape[0], self.kw.shape[0]
        w = self.module.attn_qkvw
        return w[:nq], w[nq:nq + nk], w[nq + nk:]

    def refresh_fused_qkv(self):
        """Training changed q/k/v: re-pack them into the fused layer (generation phase of RLHF)."""
        self.set_q_k_v()
        with torch.no_grad():
            self.module.attn_qkvw.copy_(torch.cat([self.qw, self.kw, self.vw], 0).to(self.module.attn_qkvw.dtype))
            if self.qb is not None:
                self.module.attn_qkvb.copy_(torch.cat([self.qb, self.kb, self.vb], 0).to(self.module.attn_qkvb.dtype))
