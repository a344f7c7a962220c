"""DeepSpeedLlama2Inference: Llama-family decoder layer (RMSNorm, rotary, gated SiLU MLP, GQA).

Reference ``model_implementations/transformers/ds_llama2.py``.  All families share one fused layer implementation
(``ops/transformer/inference/ds_transformer.py``); the family is expressed through ``DeepSpeedInferenceConfig`` fields
(pre/post layer norm, rotary dim, activation, ALiBi, ...) that the injection policy fills in."""
from deepspeed_b200.ops.transformer.inference.ds_transformer import DeepSpeedTransformerInference


class DeepSpeedLlama2Inference(DeepSpeedTransformerInference):

    def __init__(self, config, mp_group=None, quantize_scales=None, quantize_groups=1, merge_count=1, mlp_extra_grouping=False):
        super().__init__(config, mp_group, quantize_scales, quantize_groups, merge_count, mlp_extra_grouping)
