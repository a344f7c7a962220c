"""CUDA-graphed VAE wrapper (reference ``model_implementations/diffusers/vae.py``): separate graphs for ``encode``,
``decode`` and the full ``forward``."""
import torch

from ..features.cuda_graph import CUDAGraph, GraphedCallable


class DSVAE(CUDAGraph, torch.nn.Module):

    def __init__(self, vae, enable_cuda_graph=True):
        super().__init__(enable_cuda_graph=enable_cuda_graph)
        self.vae = vae
        self.config = getattr(vae, "config", None)
        self.device = getattr(vae, "device", None)
        self.dtype = getattr(vae, "dtype", None)
        self.vae.requires_grad_(False)
        self._dec = GraphedCallable(self._decode, enabled=enable_cuda_graph)
        self._enc = GraphedCallable(self._encode, enabled=enable_cuda_graph)
        self._fwd = GraphedCallable(self._forward, enabled=enable_cuda_graph)

    def _decode(self, x, return_dict=True, generator=None):
        return self.vae.decode(x, return_dict=return_dict)

    def _encode(self, x, return_dict=True):
        return self.vae.encode(x, return_dict=return_dict)

    def _forward(self, sample, timestamp=None, encoder_hidden_states=None, return_dict=True):
        return self.vae(sample, return_dict=return_dict) if timestamp is None else self.vae(
            sample, timestamp, encoder_hidden_states, return_dict)

    def decode(self, *inputs, **kwargs):
        return self._dec(*inputs, **kwargs) if self.enable_cuda_graph else self._decode(*inputs, **kwargs)

    def encode(self, *inputs, **kwargs):
        return self._enc(*inputs, **kwargs) if self.enable_cuda_graph else self._encode(*inputs, **kwargs)

    def forward(self, *inputs, **kwargs):
        return self._fwd(*inputs, **kwargs) if self.enable_cuda_graph else self._forward(*inputs, **kwargs)

    def _create_cuda_graph(self, *inputs, **kwargs):
        return self._fwd(*inputs, **kwargs)

    def _graph_replay(self, *inputs, **kwargs):
        return self._fwd(*inputs, **kwargs)
