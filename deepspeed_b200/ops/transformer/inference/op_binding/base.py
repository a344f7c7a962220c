"""Base of the v1 inference op bindings (reference ``ops/transformer/inference/op_binding/base.py``).

The reference resolves each op to a ``*_fp16`` / ``*_bf16`` / ``*_fp32`` symbol of one monolithic extension.  Here every
op is a thin callable over the typed sm_100a kernels in ``ops/kernels`` (dtype dispatch happens in the C ABI), with a
torch path for host tensors, so a binding holds nothing but the layer config.
"""
import torch

from ..config import DeepSpeedInferenceConfig


class BaseOp(torch.nn.Module):

    def __init__(self, config: DeepSpeedInferenceConfig = None):
        super().__init__()
        self.config = config if config is not None else DeepSpeedInferenceConfig()

    @property
    def eps(self):
        return getattr(self.config, "epsilon", 1e-5)
