"""Policy for OPT (LayerNorm, learned positions offset by 2, ReLU) (reference ``model_implementations/opt/policy.py``)."""
from ..inference_policy_base import ContainerMap, InferenceV2Policy
from .container import OPTNonTransformerContainer, OPTTransformerContainer
from .model import OPTInferenceModel


class OPTPolicy(InferenceV2Policy):
    model_type = "opt"

    def instantiate_model(self, engine_config, mp_group=None) -> OPTInferenceModel:
        import torch
        from deepspeed_b200 import comm as dist
        tp = getattr(getattr(engine_config, "tensor_parallel", None), "tp_size", 1) if engine_config is not None else 1
        rank = dist.get_rank(mp_group) if (mp_group is not None and tp > 1) else 0
        return OPTInferenceModel.from_hf_config(self._model_config, mp_group, tp, rank)

    def build_container_map(self, model=None) -> ContainerMap:
        """Declarative checkpoint map: one transformer container per layer + the non-transformer container."""
        model = model if model is not None else self.instantiate_model(None)
        cmap = ContainerMap()
        cmap.set_transformer_params(["model.decoder.layers"], [OPTTransformerContainer(model) for _ in range(model.num_layers)])
        cmap.set_non_transformer_params(OPTNonTransformerContainer(model))
        cmap.set_unmapped_params([])
        return cmap
