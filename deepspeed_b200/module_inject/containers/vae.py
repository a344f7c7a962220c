"""Diffusers VAE policy (reference ``module_inject/containers/vae.py``)."""
from ..policy import DSPolicy


class VAEPolicy(DSPolicy):

    def __init__(self):
        super().__init__()
        try:
            import diffusers
            self._orig_layer_class = getattr(diffusers.models, "AutoencoderKL", None)
        except Exception:
            self._orig_layer_class = None

    def match(self, module):
        return (self._orig_layer_class is not None and isinstance(module, self._orig_layer_class)) or \
            type(module).__name__ == "AutoencoderKL"

    def match_replaced(self, module):
        from deepspeed_b200.model_implementations.diffusers.vae import DSVAE
        return isinstance(module, DSVAE)

    def apply(self, module, enable_cuda_graph=True):
        from deepspeed_b200.model_implementations.diffusers.vae import DSVAE
        return DSVAE(module, enable_cuda_graph=enable_cuda_graph)

    def attention(self, client_module=None):
        return None
