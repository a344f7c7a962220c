"""Policy for Qwen2-MoE (routed experts + a gated shared expert) (reference ``model_implementations/qwen_v2_moe/policy.py``)."""
from ..inference_policy_base import ContainerMap, InferenceV2Policy
from .container import Qwen2MoeNonTransformerContainer, Qwen2MoeTransformerContainer
from .model import Qwen2MoeInferenceModel


class Qwen2MoePolicy(InferenceV2Policy):
    model_type = "qwen2_moe"

    def instantiate_model(self, engine_config, mp_group=None) -> Qwen2MoeInferenceModel:
        import torch
        from deepspeed_b200 import comm as dist
        tp = getattr(getattr(engine_config, "tensor_parallel", None), "tp_size", 1) if engine_config is not None else 1
        rank = dist.get_rank(mp_group) if (mp_group is not None and tp > 1) else 0
        return Qwen2MoeInferenceModel.from_hf_config(self._model_config, mp_group, tp, rank)

    def build_container_map(self, model=None) -> ContainerMap:
        """Declarative checkpoint map: one transformer container per layer + the non-transformer container."""
        model = model if model is not None else self.instantiate_model(None)
        cmap = ContainerMap()
        cmap.set_transformer_params(["model.layers"], [Qwen2MoeTransformerContainer(model) for _ in range(model.num_layers)])
        cmap.set_non_transformer_params(Qwen2MoeNonTransformerContainer(model))
        cmap.set_unmapped_params([])
        return cmap
