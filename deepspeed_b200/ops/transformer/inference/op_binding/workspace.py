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


class InferenceContext:
    """Process-wide decode state for the v1 kernel-injected layers (reference ``workspace.py:14``): the running token
    count and one ``(key, value)`` cache pair per layer shaped ``[batch, heads, max_tokens, head_dim]``.

    ``update_cache`` writes the new keys / values (whole prompt, or one decode position) and returns views of the valid
    prefix -- what an attention kernel reads."""
    _instance = None

    def __init__(self):
        self.kv_cache = None
        self.kv_cache_size = None
        self.kv_cache_elem_dtype = None
        self.num_tokens = 1
        self.static_shapes = False
        self._spec = None

    @classmethod
    def Instance(cls):
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    # ---- workspace
    def gen_workspace(self, num_layers, num_heads, batch_size, prompt_len, hidden_dim, mp_size, external_cache, elem_dtype, rank,
                      max_out_tokens, min_out_tokens):
        """Size the caches: ``max_out_tokens`` positions (at least the prompt + ``min_out_tokens``)."""
        heads = num_heads // mp_size
        tokens = max(int(max_out_tokens), int(prompt_len) + int(min_out_tokens))
        self._spec = (num_layers, batch_size, heads, tokens, hidden_dim // num_heads, elem_dtype, bool(external_cache))
        self.kv_cache = None
        self.kv_cache_elem_dtype = elem_dtype
        return self._retake_workspace()

    def retake_workspace(self):
        return True

    def _retake_workspace(self):
        assert self._spec is not None, "Need to call gen_workspace once to set up the workspace"
        if self.kv_cache is None:
            layers, b, h, t, d, dt, external = self._spec
            self.kv_cache_size = (b, h, t, d)
            if external:
                self.kv_cache = []
                return True
            dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
            self.kv_cache = [(torch.zeros(self.kv_cache_size, dtype=dt, device=dev),
                              torch.zeros(self.kv_cache_size, dtype=dt, device=dev)) for _ in range(layers)]
        return True

    def release_workspace(self):
        self.kv_cache = None

    # ---- token counter
    def reset_tokens(self, initial_tokens=1):
        self.num_tokens = initial_tokens

    def current_tokens(self):
        return self.num_tokens

    def advance_tokens(self):
        self.num_tokens = self.num_tokens + 1

    # ---- cache
    def update_cache(self, layer_id, token_idx, is_prompt, bat_0213_key, bat_0213_value):
        """``bat_0213_*``: ``[batch, heads, seq, head_dim]``.  Prompt: cache[:, :, :seq] = k/v and the tail is zeroed;
        decode: position ``token_idx-1`` is written (``token_idx`` may be a 1-element tensor for static-shape graphs)."""
        assert self._retake_workspace(), "Could not allocate workspace"
        if is_prompt:
            self.static_shapes = token_idx is not None
            self.reset_tokens(bat_0213_key.shape[2] if token_idx is None else token_idx)
        if token_idx is None:
            token_idx = self.current_tokens()
        b = bat_0213_key.shape[0]
        kc, vc = self.kv_cache[layer_id]
        if is_prompt:
            seq = bat_0213_key.shape[2]
            for cache, new in ((kc, bat_0213_key), (vc, bat_0213_value)):
                cache[:b, :, :seq].copy_(new)
                dead = torch.arange(cache.shape[2], device=cache.device) >= (token_idx if torch.is_tensor(token_idx) else int(token_idx))
                cache[:b].masked_fill_(dead.view(1, 1, -1, 1), 0)
        elif self.static_shapes:
            assert torch.is_tensor(token_idx), "token_idx is expected to be torch.Tensor"
            kc[:b].index_copy_(2, token_idx.reshape(-1) - 1, bat_0213_key)
            vc[:b].index_copy_(2, token_idx.reshape(-1) - 1, bat_0213_value)
        else:
            assert isinstance(token_idx, int), "token_idx is expected to be int"
            kc[:b, :, token_idx - 1:token_idx] = bat_0213_key
            vc[:b, :, token_idx - 1:token_idx] = bat_0213_value
        k, v = kc[:b], vc[:b]
        if not self.static_shapes:
            k, v = k[:, :, :token_idx], v[:, :, :token_idx]
        return k, v
