"""CUDA-graphed UNet wrapper for Stable-Diffusion pipelines (reference ``model_implementations/diffusers/unet.py``)."""
import torch

from ..features.cuda_graph import CUDAGraph, GraphedCallable


class DSUNet(CUDAGraph, torch.nn.Module):

    def __init__(self, unet, enable_cuda_graph=True):
        super().__init__(enable_cuda_graph=enable_cuda_graph)
        self.unet = unet
        # attributes the diffusers pipelines read from the UNet
        self.in_channels = getattr(unet, "in_channels", getattr(getattr(unet, "config", None), "in_channels", None))
        self.device = getattr(unet, "device", None)
        self.dtype = getattr(unet, "dtype", None)
        self.config = getattr(unet, "config", None)
        self.fwd_count = 0
        self.unet.requires_grad_(False)
        self.unet.to(memory_format=torch.channels_last)
        self._graphed = GraphedCallable(self._forward, enabled=enable_cuda_graph)

    @property
    def cuda_graph_created(self):
        return self._graphed.captures > 0

    def _create_cuda_graph(self, *inputs, **kwargs):
        return self._graphed(*inputs, **kwargs)

    def _graph_replay(self, *inputs, **kwargs):
        return self._graphed(*inputs, **kwargs)

    def forward(self, *inputs, **kwargs):
        self.fwd_count += 1
        return self._graphed(*inputs, **kwargs) if self.enable_cuda_graph else self._forward(*inputs, **kwargs)

    def _forward(self, sample, timestamp, encoder_hidden_states, return_dict=True, cross_attention_kwargs=None,
                 timestep_cond=None, added_cond_kwargs=None):
        extra = {}
        if cross_attention_kwargs:
            extra["cross_attention_kwargs"] = cross_attention_kwargs
        if timestep_cond is not None:
            extra["timestep_cond"] = timestep_cond
        if added_cond_kwargs:
            extra["added_cond_kwargs"] = added_cond_kwargs
        return self.unet(sample, timestamp, encoder_hidden_states, return_dict=return_dict, **extra)
