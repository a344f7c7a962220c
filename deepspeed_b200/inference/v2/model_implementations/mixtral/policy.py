"""Policy for Mixtral 8x7B-style sparse MoE (top-2 of 8 gated experts) (reference ``model_implementations/mixtral/policy.py``)."""
from ..inference_policy_base import ContainerMap, InferenceV2Policy
from .container import MixtralNonTransformerContainer, MixtralTransformerContainer
from .model import MixtralInferenceModel


class MixtralPolicy(InferenceV2Policy):
    model_type = "mixtral"

    def instantiate_model(self, engine_config, mp_group=None) -> MixtralInferenceModel:
        import torch
        from deepspeed_b200 import comm as dist
        tp = getattr(getattr(engine_config, "tensor_parallel", None), "tp_size", 1) if engine_config is not None else 1
        rank = dist.get_rank(mp_group) if (mp_group is not None and tp > 1) else 0
        return MixtralInferenceModel.from_hf_config(self._model_config, mp_group, tp, rank)

    def build_container_map(self, model=None) -> ContainerMap:
        """Declarative checkpoint map: one transformer container per layer + the non-transformer container."""
        model = model if model is not None else self.instantiate_model(None)
        cmap = ContainerMap()
        cmap.set_transformer_params(["model.layers"], [MixtralTransformerContainer(model) for _ in range(model.num_layers)])
        cmap.set_non_transformer_params(MixtralNonTransformerContainer(model))
        cmap.set_unmapped_params([])
        return cmap
