from deepspeed_b200.ops.transformer.inference.ds_transformer import DeepSpeedTransformerInference  # noqa: F401
