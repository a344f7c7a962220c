"""CUDA-graphed CLIP text-encoder wrapper (reference ``model_implementations/transformers/clip_encoder.py``).  A
Stable-Diffusion step calls the encoder twice (prompt + negative prompt); the signature-keyed graph cache covers both."""
import torch

from ..features.cuda_graph import CUDAGraph, GraphedCallable


class DSClipEncoder(CUDAGraph, torch.nn.Module):

    def __init__(self, enc, enable_cuda_graph=False):
        super().__init__(enable_cuda_graph=enable_cuda_graph)
        tm = getattr(enc, "text_model", None)
        if tm is not None and hasattr(tm, "_build_causal_attention_mask"):
            tm._build_causal_attention_mask = self._build_causal_attention_mask  # allocation-free under graph capture
        self.enc = enc
        self.device = getattr(enc, "device", None)
        self.dtype = getattr(enc, "dtype", None)
        self.config = getattr(enc, "config", None)
        self._graphed = GraphedCallable(self._forward, enabled=enable_cuda_graph)

    def _build_causal_attention_mask(self, bsz, seq_len, dtype):
        dev = self.device if self.device is not None else "cpu"
        mask = torch.full((bsz, seq_len, seq_len), torch.finfo(dtype).min, dtype=dtype, device=dev)
        return mask.triu_(1).unsqueeze(1)

    def _forward(self, *inputs, **kwargs):
        return self.enc(*inputs, **kwargs)

    def forward(self, *inputs, **kwargs):
        return self._graphed(*inputs, **kwargs) if self.enable_cuda_graph else self._forward(*inputs, **kwargs)

    def _create_cuda_graph(self, *inputs, **kwargs):
        return self._graphed(*inputs, **kwargs)

    def _graph_replay(self, *inputs, **kwargs):
        return self._graphed(*inputs, **kwargs)
