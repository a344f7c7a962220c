"""MoE variant of the container (reference ``containers/base_moe.py``): expert MLPs come as lists, one per local expert."""
from .base import BaseTransformerContainer


class BaseTransformerMoEContainer(BaseTransformerContainer):

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.num_experts = getattr(self.policy, "num_experts", 1)
        self.ep_world_size = 1

    def initialize_tensors(self, enable_training=False):
        self.qkvw, self.qkvb, self.dense_w, self.dense_b = self.policy.attention()
        self.attn_nw, self.attn_nb, self.input_nw, self.input_nb = self.policy.layernorm()
        # (w1, b1, w2, b2) per expert
        self.expert_mlps = [self.policy.mlp(moe_type="standard", expert=e) if "expert" in self.policy.mlp.__code__.co_varnames
                            else self.policy.mlp() for e in range(self.num_experts)]
        self._h4h_w, self._h4h_b, self._4hh_w, self._4hh_b = self.expert_mlps[0]
