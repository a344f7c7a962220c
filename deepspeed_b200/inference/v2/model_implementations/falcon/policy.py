"""Policy for Falcon (parallel attention/MLP, multi-query or grouped fused QKV) (reference ``model_implementations/falcon/policy.py``)."""
from ..inference_policy_base import ContainerMap, InferenceV2Policy
from .container import FalconNewArchTransformerContainer, FalconNonTransformerContainer, FalconTransformerContainer
from .model import FalconInferenceModel


class FalconPolicy(InferenceV2Policy):
    model_type = "falcon"

    def instantiate_model(self, engine_config, mp_group=None) -> FalconInferenceModel:
        import torch
        from deepspeed_b200 import comm as dist
        tp = getattr(getattr(engine_config, "tensor_parallel", None), "tp_size", 1) if engine_config is not None else 1
        rank = dist.get_rank(mp_group) if (mp_group is not None and tp > 1) else 0
        return FalconInferenceModel.from_hf_config(self._model_config, mp_group, tp, rank)

    def build_container_map(self, model=None) -> ContainerMap:
        """Declarative checkpoint map: one transformer container per layer + the non-transformer container."""
        model = model if model is not None else self.instantiate_model(None)
        cmap = ContainerMap()
        new_arch = bool(getattr(model.spec, "extras", {}).get("new_decoder_architecture"))
        layer_cls = FalconNewArchTransformerContainer if new_arch else FalconTransformerContainer
        cmap.set_transformer_params(["transformer.h"], [layer_cls(model) for _ in range(model.num_layers)])
        cmap.set_non_transformer_params(FalconNonTransformerContainer(model))
        cmap.set_unmapped_params([])
        return cmap
