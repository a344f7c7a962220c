"""Diffusers UNet policy: the whole module is wrapped (CUDA graphs) and its attention blocks fused (reference
``module_inject/containers/unet.py``)."""
from ..policy import DSPolicy


def _cls(path, name):
    try:
        import importlib
        return getattr(importlib.import_module(path), name)
    except Exception:
        return None


class UNetPolicy(DSPolicy):

    def __init__(self):
        super().__init__()
        self._orig_layer_class = _cls("diffusers.models.unet_2d_condition", "UNet2DConditionModel") or \
            _cls("diffusers.models.unets.unet_2d_condition", "UNet2DConditionModel")

    def match(self, module):
        return (self._orig_layer_class is not None and isinstance(module, self._orig_layer_class)) or \
            type(module).__name__ == "UNet2DConditionModel"

    def match_replaced(self, module):
        from deepspeed_b200.model_implementations.diffusers.unet import DSUNet
        return isinstance(module, DSUNet)

    def apply(self, module, enable_cuda_graph=True):
        from deepspeed_b200.model_implementations.diffusers.unet import DSUNet
        from deepspeed_b200.module_inject.replace_module import generic_injection
        generic_injection(module, enable_cuda_graph=enable_cuda_graph)
        return DSUNet(module, enable_cuda_graph=enable_cuda_graph)

    def attention(self, client_module=None):
        a = client_module
        return (a.to_q.weight, a.to_k.weight, a.to_v.weight), None, a.to_out[0].weight, a.to_out[0].bias
