"""Mistral (llama layout, sliding-window capable) on the ragged engine (reference ``model_implementations/mistral/model.py``).

The family is a configuration of ``RaggedTransformer`` (see ``arch.arch_from_hf_config``); this class pins the model type
and provides the parameter-transform hooks the declarative containers call."""
from ..arch import arch_from_hf_config
from ..ragged_transformer import RaggedTransformer
from ..transforms import ContainerTransformsMixin


class MistralInferenceModel(ContainerTransformsMixin, RaggedTransformer):
    model_type = "mistral"

    @classmethod
    def from_hf_config(cls, hf_config, tp_group=None, tp_size=1, tp_rank=0, dtype=None, device=None):
        import torch
        spec = arch_from_hf_config(hf_config)
        assert spec.model_type == cls.model_type, f"{cls.__name__} cannot serve model type {spec.model_type}"
        return cls(spec, tp_group, tp_size, tp_rank, dtype or torch.bfloat16, device or ("cuda" if torch.cuda.is_available() else "cpu"))
