class Diffusers2DTransformerConfig:
    """Config of the fused diffusers transformer block (reference ``diffusers_2d_transformer.py``)."""

    def __init__(self, int8_quantization=False):
        self.int8_quantization = int8_quantization
