"""DeepSpeedOPTInference: OPT decoder layer (learned positions, ReLU MLP).

Reference ``model_implementations/transformers/ds_opt.py``.  All families share one fused layer implementation
(``ops/transformer/inference/ds_transformer.py``); the family is expressed through ``DeepSpeedInferenceConfig`` fields
(pre/post layer norm, rotary dim, activation, ALiBi, ...) that the injection policy fills in."""
from deepspeed_b200.ops.transformer.inference.ds_transformer import DeepSpeedTransformerInference


class DeepSpeedOPTInference(DeepSpeedTransformerInference):

    def __init__(self, config, mp_group=None, quantize_scales=None, quantize_groups=1, merge_count=1, mlp_extra_grouping=False):
        super().__init__(config, mp_group, quantize_scales, quantize_groups, merge_count, mlp_extra_grouping)
