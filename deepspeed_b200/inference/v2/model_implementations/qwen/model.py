"""Qwen-1 (fused c_attn with biases, w1/w2 gated MLP) on the ragged engine (reference ``model_implementations/qwen/model.py``).

The family is a configuration of ``RaggedTransformer`` (see ``arch.arch_from_hf_config``); this class pins the model type
and provides the parameter-transform hooks the declarative containers call."""
from ..arch import arch_from_hf_config
from ..ragged_transformer import RaggedTransformer
from ..transforms import ContainerTransformsMixin


class QwenInferenceModel(ContainerTransformsMixin, RaggedTransformer):
    model_type = "qwen"

    @classmethod
    def from_hf_config(cls, hf_config, tp_group=None, tp_size=1, tp_rank=0, dtype=None, device=None):
        import torch
        spec = arch_from_hf_config(hf_config)
        assert spec.model_type == cls.model_type, f"{cls.__name__} cannot serve model type {spec.model_type}"
        return cls(spec, tp_group, tp_size, tp_rank, dtype or torch.bfloat16, device or ("cuda" if torch.cuda.is_available() else "cpu"))
