"""``WorkspaceOp`` (reference ``ops/transformer/inference/op_binding/workspace.py``): the per-process inference workspace.

The reference carves activations and every layer's KV cache out of one device arena sized at first use.  Here
activations come from the caching allocator (CUDA-graph friendly already) and the workspace owns what must persist
across calls: one ``[batch, 2, kv_heads, max_out_tokens, head_dim]`` KV cache per layer plus its fill level.
"""
import torch

from .base import BaseOp


class WorkspaceOp(BaseOp):
    _caches = {}
    _seen = {}
    _allocated = False

    def allocate_workspace(self, hidden_dim, num_heads, prompt_length, batch_size, num_layers, mp_size=1, external_cache=False,
                           rank=0, max_out_tokens=1024, min_out_tokens=1):
        WorkspaceOp._allocated = True
        self.config.max_out_tokens = max(int(max_out_tokens), int(min_out_tokens))
        return True

    def is_allocated(self):
        return WorkspaceOp._allocated

    def release_workspace(self):
        WorkspaceOp._caches.clear()
        WorkspaceOp._seen.clear()
        WorkspaceOp._allocated = False
        return True

    def retake_workspace(self):
        WorkspaceOp._allocated = True
        return True

    def reset_cache(self):
        WorkspaceOp._seen.clear()

    @classmethod
    def kv_cache(cls, layer_id, batch, kv_heads, max_tokens, head_dim, like):
        c = cls._caches.get(layer_id)
        if c is None or c.shape[0] < batch or c.shape[3] < max_tokens or c.device != like.device or c.dtype != like.dtype:
            c = cls._caches[layer_id] = torch.zeros(batch, 2, kv_heads, max_tokens, head_dim, dtype=like.dtype, device=like.device)
        return c

    @classmethod
    def seen(cls, layer_id):
        return cls._seen.get(layer_id, 0)

    @classmethod
    def set_seen(cls, layer_id, n):
        cls._seen[layer_id] = n
