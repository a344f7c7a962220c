"""Host-side builder of the ragged batch metadata (reference ``ragged/ragged_wrapper.py:31``).

All metadata is staged in ONE pinned host buffer and moved with a single async H2D copy into a persistent
device buffer (so the decode step is CUDA-graph friendly: the graph reads fixed addresses):

    input_ids [T] | seq_of [T] | pos_of [T] | last_tok [S] | block_table [S, max_blocks]
"""
import numpy as np
import torch

from deepspeed_b200.accelerator import get_accelerator
from ..config_v2 import DSStateManagerConfig


class RaggedBatchWrapper:

    def __init__(self, config: DSStateManagerConfig, max_blocks_per_seq: int = 64, device=None):
        self._config = config
        T, S, MB = config.max_ragged_batch_size, config.max_ragged_sequence_count, max_blocks_per_seq
        self._max_T, self._max_S, self._max_blocks = T, S, MB
        self.device = device if device is not None else get_accelerator().current_device_name()
        total = 3 * T + S + S * MB
        pin = torch.cuda.is_available()
        self._host = torch.zeros(total, dtype=torch.int32, pin_memory=pin)
        self._dev = torch.zeros(total, dtype=torch.int32, device=self.device)
        o = 0
        self._views = {}
        for name, n in (("ids", T), ("seq_of", T), ("pos_of", T), ("last_tok", S), ("block_table", S * MB)):
            self._views[name] = (o, n)
            o += n
        # numpy views of the pinned staging buffer: per-sequence bookkeeping is pure host work on the critical path of
        # every decode step, and numpy slice writes cost ~10x less than torch indexing
        hn = self._host.numpy()
        self._np = {k: hn[o:o + n] for k, (o, n) in self._views.items()}
        self._np["block_table"] = self._np["block_table"].reshape(S, MB)
        self.clear()

    def _h(self, name):
        o, n = self._views[name]
        return self._host[o:o + n]

    def _d(self, name):
        o, n = self._views[name]
        return self._dev[o:o + n]

    def clear(self) -> None:
        self._n_tokens = 0
        self._n_seqs = 0
        self._seq_tokens = []
        self._seq_seen = []
        self._is_finalized = False

    def insert_sequence(self, seq_descriptor, tokens, do_checks=True) -> None:
        tok = tokens.numpy() if isinstance(tokens, torch.Tensor) else tokens
        n = tok.size if hasattr(tok, "size") else len(tok)
        if do_checks:
            if self._n_seqs + 1 > self._max_S:
                raise RuntimeError(f"Ragged batch is full: {self._n_seqs} sequences")
            if self._n_tokens + n > self._max_T:
                raise RuntimeError(f"Ragged batch is full: {self._n_tokens} + {n} tokens > {self._max_T}")
        s, t0 = self._n_seqs, self._n_tokens
        seen = seq_descriptor.seen_tokens
        v = self._np
        if n == 1:
            v["ids"][t0] = tok.reshape(-1)[0] if hasattr(tok, "reshape") else tok[0]
            v["seq_of"][t0] = s
            v["pos_of"][t0] = seen
        else:
            v["ids"][t0:t0 + n] = tok.reshape(-1) if hasattr(tok, "reshape") else tok
            v["seq_of"][t0:t0 + n] = s
            v["pos_of"][t0:t0 + n] = np.arange(seen, seen + n, dtype=np.int32)
        v["last_tok"][s] = t0 + n - 1
        nb = seq_descriptor.cur_allocated_blocks
        v["block_table"][s, :nb] = seq_descriptor.kv_ids_np(0)[:nb]
        self._seq_tokens.append(n)
        self._seq_seen.append(seen)
        self._n_seqs += 1
        self._n_tokens += n

    def finalize(self, padding: bool = False) -> None:
        """One H2D copy of the used prefix of every region (regions are contiguous so a single copy of the whole
        staging buffer is cheaper than five small ones at these sizes)."""
        self._dev.copy_(self._host, non_blocking=True)
        self._is_finalized = True

    # --- views consumed by the model ---
    def input_ids(self, padded_tokens=None):
        return self._d("ids")[:padded_tokens or self._n_tokens]

    def seq_of(self, padded_tokens=None):
        return self._d("seq_of")[:padded_tokens or self._n_tokens]

    def pos_of(self, padded_tokens=None):
        return self._d("pos_of")[:padded_tokens or self._n_tokens]

    def last_token_index(self, padded_seqs=None):
        return self._d("last_tok")[:padded_seqs or self._n_seqs]

    def block_table(self):
        return self._d("block_table").view(self._max_S, self._max_blocks)

    @property
    def seq_layout(self):
        """Host list of (token_start, n_tokens, seen_tokens) per sequence — used to pick dense prefill."""
        out, t = [], 0
        for n, seen in zip(self._seq_tokens, self._seq_seen):
            out.append((t, n, seen))
            t += n
        return out

    @property
    def current_tokens(self) -> int:
        return self._n_tokens

    @property
    def current_sequences(self) -> int:
        return self._n_seqs

    @property
    def tensor_toks(self):
        return self._n_tokens

    @property
    def is_pure_decode(self) -> bool:
        return self._n_seqs == self._n_tokens and self._n_seqs > 0

    # ---- reference-named views (``ragged_wrapper.py:230-280``) -------------------------------------------------------
    def batch_metadata_buffer(self, on_device: bool = True) -> torch.Tensor:
        """``[n_tokens, n_sequences]`` as int32 (what the reference's kernels read first)."""
        t = torch.tensor([self._n_tokens, self._n_seqs], dtype=torch.int32)
        return t.to(self.device) if on_device else t

    def tokens_to_seq(self, on_device: bool = True) -> torch.Tensor:
        """Sequence slot of every token in the batch."""
        return self.seq_of() if on_device else self._h("seq_of")[:self._n_tokens]

    def inflight_seq_descriptors(self, on_device: bool = True) -> torch.Tensor:
        """``[n_sequences, 4]`` rows of (first token, tokens in this batch, tokens already seen, 0)."""
        rows = [(t0, n, seen, 0) for (t0, n, seen) in self.seq_layout]
        t = torch.tensor(rows, dtype=torch.int32).reshape(-1, 4)
        return t.to(self.device) if on_device else t

    def kv_ptrs(self, on_device: bool = True) -> torch.Tensor:
        """Per-sequence KV block ids (the block table plays the role of the reference's pointer array)."""
        bt = self.block_table() if on_device else self._h("block_table").view(self._max_S, self._max_blocks)
        return bt[:self._n_seqs]

    def masks(self, on_device: bool = True):
        """No explicit masks: causal masking is derived from token positions inside the attention kernels."""
        return None


def to_padded(original_size: int) -> int:
    """Round a token / sequence count up to the granularity CUDA-graph buckets and GEMM tiles like: 64 up to 512, 128
    above (reference ``ragged_wrapper.py:17``)."""
    g = 64 if original_size <= 512 else 128
    return (original_size + g - 1) // g * g
