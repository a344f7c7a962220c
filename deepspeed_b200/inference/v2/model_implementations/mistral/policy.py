"""Policy for Mistral (llama layout, sliding-window capable) (reference ``model_implementations/mistral/policy.py``)."""
from ..inference_policy_base import ContainerMap, InferenceV2Policy
from .container import MistralNonTransformerContainer, MistralTransformerContainer
from .model import MistralInferenceModel


class MistralPolicy(InferenceV2Policy):
    model_type = "mistral"

    def instantiate_model(self, engine_config, mp_group=None) -> MistralInferenceModel:
        import torch
        from deepspeed_b200 import comm as dist
        tp = getattr(getattr(engine_config, "tensor_parallel", None), "tp_size", 1) if engine_config is not None else 1
        rank = dist.get_rank(mp_group) if (mp_group is not None and tp > 1) else 0
        return MistralInferenceModel.from_hf_config(self._model_config, mp_group, tp, rank)

    def build_container_map(self, model=None) -> ContainerMap:
        """Declarative checkpoint map: one transformer container per layer + the non-transformer container."""
        model = model if model is not None else self.instantiate_model(None)
        cmap = ContainerMap()
        cmap.set_transformer_params(["model.layers"], [MistralTransformerContainer(model) for _ in range(model.num_layers)])
        cmap.set_non_transformer_params(MistralNonTransformerContainer(model))
        cmap.set_unmapped_params([])
        return cmap
