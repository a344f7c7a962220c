"""Block-sparse attention layouts.  Behavioural parity: reference ``ops/sparse_attention/sparsity_config.py``
(Dense / Fixed / Variable / BigBird / BSLongformer / LocalSlidingWindow).  ``make_layout(seq_len)`` returns an
int64 ``[num_heads, seq/block, seq/block]`` 0/1 tensor; layouts are built with vectorised index arithmetic."""
import random

import torch


class SparsityConfig:

    def __init__(self, num_heads, block=16, different_layout_per_head=False):
        self.num_heads = num_heads
        self.block = block
        self.different_layout_per_head = different_layout_per_head
        self.num_layout_heads = num_heads if different_layout_per_head else 1

    def setup_layout(self, seq_len):
        if seq_len % self.block != 0:
            raise ValueError(f"Sequence Length, {seq_len}, needs to be dividable by Block size {self.block}!")
        n = seq_len // self.block
        return torch.zeros((self.num_heads, n, n), dtype=torch.int64)

    def check_and_propagate_first_head_layout(self, layout):
        if not self.different_layout_per_head:
            layout[1:self.num_heads] = layout[0]
        return layout

    def make_layout(self, seq_len):
        raise NotImplementedError


class DenseSparsityConfig(SparsityConfig):

    def make_layout(self, seq_len):
        layout = self.setup_layout(seq_len)
        layout[:] = 1
        return layout


def _grid(n):
    i = torch.arange(n)
    return i[:, None], i[None, :]


class FixedSparsityConfig(SparsityConfig):
    """Local windows of ``num_local_blocks`` + the last ``num_global_blocks`` of every window as global columns
    (Sparse Transformers 'fixed' pattern)."""

    def __init__(self, num_heads, block=16, different_layout_per_head=False, num_local_blocks=4, num_global_blocks=1,
                 attention="bidirectional", horizontal_global_attention=False, num_different_global_patterns=1):
        super().__init__(num_heads, block, different_layout_per_head)
        self.num_local_blocks = num_local_blocks
        if num_local_blocks % num_global_blocks != 0:
            raise ValueError(f"Number of blocks in a local window, {num_local_blocks}, must be dividable by number of "
                             f"global blocks, {num_global_blocks}!")
        self.num_global_blocks = num_global_blocks
        if attention not in ("unidirectional", "bidirectional"):
            raise NotImplementedError('only "uni/bi-directional" attentions are supported for now!')
        self.attention = attention
        if attention != "bidirectional" and horizontal_global_attention:
            raise ValueError('only "bi-directional" attentions can support horizontal global attention!')
        self.horizontal_global_attention = horizontal_global_attention
        if num_different_global_patterns > 1 and not different_layout_per_head:
            raise ValueError("Number of different layouts cannot be more than one when you have set a single layout "
                             "for all heads! Set different_layout_per_head to True.")
        if num_different_global_patterns > (num_local_blocks // num_global_blocks):
            raise ValueError(f"Number of layout versions (num_different_global_patterns), "
                             f"{num_different_global_patterns}, cannot be larger than number of local window blocks "
                             f"divided by number of global blocks, {num_local_blocks} / {num_global_blocks} = "
                             f"{num_local_blocks // num_global_blocks}!")
        self.num_different_global_patterns = num_different_global_patterns

    def make_layout(self, seq_len):
        layout = self.setup_layout(seq_len)
        n = layout.shape[1]
        r, c = _grid(n)
        L, G = self.num_local_blocks, self.num_global_blocks
        local = (r // L == c // L)
        if self.attention == "unidirectional":
            local &= c <= r
        for h in range(self.num_layout_heads):
            m = local.clone()
            first = L - (1 + h % self.num_different_global_patterns) * G
            # global columns: [first, first+G) inside every full window; the tail window uses its last G blocks
            end = n - n % L
            cols = torch.zeros(n, dtype=torch.bool)
            for s in range(first, end, L):
                cols[s:s + G] = True
            if end < n:
                s = min(end + first, n - G)
                cols[s:s + G] = True
            gm = cols[None, :].expand(n, n).clone()
            if self.attention == "unidirectional":
                # a row sees a global column only from the window after it
                first_row = (torch.arange(n)[None, :] // L) * L  # window start of the column
                gm &= r >= torch.minimum(first_row, torch.tensor(n - 1))
                gm &= c <= r
            m |= gm
            if self.horizontal_global_attention:
                m |= cols[:, None].expand(n, n)
            layout[h] = m.long()
        return self.check_and_propagate_first_head_layout(layout)

    def set_local_layout(self, h, layout):
        """OR the block-diagonal local windows (``num_local_blocks`` wide; lower triangle only when causal) into head ``h``."""
        n = layout.shape[1]
        r, c = _grid(n)
        m = (r // self.num_local_blocks == c // self.num_local_blocks)
        if self.attention == "unidirectional":
            m &= c <= r
        layout[h] |= m.long()
        return layout

    def set_global_layout(self, h, layout):
        """OR the global columns of head ``h`` (last ``num_global_blocks`` of every window, rotated per head when
        ``num_different_global_patterns`` > 1) -- and the matching rows with ``horizontal_global_attention``."""
        n = layout.shape[1]
        L, G = self.num_local_blocks, self.num_global_blocks
        r, c = _grid(n)
        first = L - (1 + h % self.num_different_global_patterns) * G
        end = n - n % L
        cols = torch.zeros(n, dtype=torch.bool)
        for s0 in range(first, end, L):
            cols[s0:s0 + G] = True
        if end < n:
            s0 = min(end + first, n - G)
            cols[s0:s0 + G] = True
        gm = cols[None, :].expand(n, n).clone()
        if self.attention == "unidirectional":
            gm &= c <= r
        layout[h] |= gm.long()
        if self.horizontal_global_attention:
            layout[h] |= cols[:, None].expand(n, n).long()
        return layout


class VariableSparsityConfig(SparsityConfig):
    """Random blocks + variable-size local windows + explicit global block indices."""

    def __init__(self, num_heads, block=16, different_layout_per_head=False, num_random_blocks=0, local_window_blocks=[4],
                 global_block_indices=[0], global_block_end_indices=None, attention="bidirectional",
                 horizontal_global_attention=False):
        super().__init__(num_heads, block, different_layout_per_head)
        self.num_random_blocks = num_random_blocks
        self.local_window_blocks = local_window_blocks
        self.global_block_indices = global_block_indices
        if global_block_end_indices is not None:
            if len(global_block_indices) != len(global_block_end_indices):
                raise ValueError(f"Global block start indices length, {len(global_block_indices)}, must be same as "
                                 f"global block end indices length, {len(global_block_end_indices)}!")
            for s, e in zip(global_block_indices, global_block_end_indices):
                if s >= e:
                    raise ValueError(f"Global block start index, {s}, must be smaller than global block end index, {e}!")
        self.global_block_end_indices = global_block_end_indices
        if attention not in ("unidirectional", "bidirectional"):
            raise NotImplementedError('only "uni/bi-directional" attentions are supported for now!')
        self.attention = attention
        if attention != "bidirectional" and horizontal_global_attention:
            raise ValueError('only "bi-directional" attentions can support horizontal global attention!')
        self.horizontal_global_attention = horizontal_global_attention

    def make_layout(self, seq_len):
        layout = self.setup_layout(seq_len)
        n = layout.shape[1]
        r, c = _grid(n)
        uni = self.attention == "unidirectional"
        for h in range(self.num_layout_heads):
            m = torch.zeros(n, n, dtype=torch.bool)
            if self.num_random_blocks:
                if n < self.num_random_blocks:
                    raise ValueError(f"Number of random blocks, {self.num_random_blocks}, must be smaller than overall "
                                     f"number of blocks in a row, {n}!")
                for row in range(n):
                    hi = n if not uni else row + 1
                    for col in random.sample(range(hi), min(self.num_random_blocks, hi)):
                        m[row, col] = True
            start = 0
            sizes = list(self.local_window_blocks)
            while start < n:
                w = sizes.pop(0) if sizes else self.local_window_blocks[-1]
                end = min(start + w, n)
                blk = (r >= start) & (r < end) & (c >= start) & (c < end)
                if uni:
                    blk &= c <= r
                m |= blk
                start = end
            ends = self.global_block_end_indices or [i + 1 for i in self.global_block_indices]
            for s, e in zip(self.global_block_indices, ends):
                if s < n:
                    e = min(e, n)
                    col = (c >= s) & (c < e)
                    m |= (col & (r >= s)) if uni else col.expand(n, n)
                    if self.horizontal_global_attention:
                        m |= ((r >= s) & (r < e)).expand(n, n)
            layout[h] = m.long()
        return self.check_and_propagate_first_head_layout(layout)

    def set_random_layout(self, h, layout):
        """OR ``num_random_blocks`` random block columns per block row into head ``h`` (causal rows draw from their past)."""
        n = layout.shape[1]
        if n < self.num_random_blocks:
            raise ValueError(f"Number of random blocks, {self.num_random_blocks}, must be smaller than overall number of "
                             f"blocks in a row, {n}!")
        uni = getattr(self, "attention", "bidirectional") == "unidirectional"
        for row in range(n):
            hi = n if not uni else row + 1
            for col in random.sample(range(hi), min(self.num_random_blocks, hi)):
                layout[h, row, col] = 1
        return layout

    def set_local_layout(self, h, layout):
        """OR the variable-width local windows (``local_window_blocks``; the last width repeats) into head ``h``."""
        n = layout.shape[1]
        r, c = _grid(n)
        uni = self.attention == "unidirectional"
        start, sizes = 0, list(self.local_window_blocks)
        while start < n:
            w = sizes.pop(0) if sizes else self.local_window_blocks[-1]
            end = min(start + w, n)
            blk = (r >= start) & (r < end) & (c >= start) & (c < end)
            if uni:
                blk &= c <= r
            layout[h] |= blk.long()
            start = end
        return layout

    def set_global_layout(self, h, layout):
        """OR the global block columns ``[start, end)`` (rows too with ``horizontal_global_attention``) into head ``h``."""
        n = layout.shape[1]
        r, c = _grid(n)
        uni = self.attention == "unidirectional"
        ends = self.global_block_end_indices or [i + 1 for i in self.global_block_indices]
        for s0, e in zip(self.global_block_indices, ends):
            if s0 < n:
                e = min(e, n)
                col = (c >= s0) & (c < e)
                layout[h] |= ((col & (r >= s0)) if uni else col.expand(n, n)).long()
                if self.horizontal_global_attention:
                    layout[h] |= ((r >= s0) & (r < e)).expand(n, n).long()
        return layout


class BigBirdSparsityConfig(SparsityConfig):

    def __init__(self, num_heads, block=16, different_layout_per_head=False, num_random_blocks=1,
                 num_sliding_window_blocks=3, num_global_blocks=1, attention="bidirectional"):
        super().__init__(num_heads, block, different_layout_per_head)
        self.num_random_blocks = num_random_blocks
        self.num_sliding_window_blocks = num_sliding_window_blocks
        self.num_global_blocks = num_global_blocks
        if attention not in ("unidirectional", "bidirectional"):
            raise NotImplementedError('only "uni/bi-directional" attentions are supported for now!')
        self.attention = attention

    def make_layout(self, seq_len):
        layout = self.setup_layout(seq_len)
        n = layout.shape[1]
        for k, name in ((self.num_random_blocks, "random"), (self.num_sliding_window_blocks, "sliding window"),
                        (self.num_global_blocks, "global")):
            if n < k:
                raise ValueError(f"Number of {name} blocks, {k}, must be smaller than overall number of blocks in a row, "
                                 f"{n}!")
        r, c = _grid(n)
        uni = self.attention == "unidirectional"
        w = self.num_sliding_window_blocks // 2
        for h in range(self.num_layout_heads):
            m = (c >= r - w) & (c <= r + w)
            for row in range(n):
                hi = n if not uni else row + 1
                for col in random.sample(range(hi), min(self.num_random_blocks, hi)):
                    m[row, col] = True
            g = self.num_global_blocks
            m |= (c < g).expand(n, n)
            m |= (r < g).expand(n, n)
            if uni:
                m &= c <= r
            layout[h] = m.long()
        return self.check_and_propagate_first_head_layout(layout)

    def set_random_layout(self, h, layout):
        """OR ``num_random_blocks`` random block columns per block row into head ``h`` (causal rows draw from their past)."""
        n = layout.shape[1]
        if n < self.num_random_blocks:
            raise ValueError(f"Number of random blocks, {self.num_random_blocks}, must be smaller than overall number of "
                             f"blocks in a row, {n}!")
        uni = getattr(self, "attention", "bidirectional") == "unidirectional"
        for row in range(n):
            hi = n if not uni else row + 1
            for col in random.sample(range(hi), min(self.num_random_blocks, hi)):
                layout[h, row, col] = 1
        return layout

    def set_sliding_window_layout(self, h, layout):
        """OR the band ``|row - col| <= num_sliding_window_blocks // 2`` into head ``h``."""
        n = layout.shape[1]
        if n < self.num_sliding_window_blocks:
            raise ValueError(f"Number of sliding window blocks, {self.num_sliding_window_blocks}, must be smaller than "
                             f"overall number of blocks in a row, {n}!")
        r, c = _grid(n)
        w = self.num_sliding_window_blocks // 2
        layout[h] |= ((c >= r - w) & (c <= r + w)).long()
        return layout

    def set_global_layout_itc(self, h, layout):
        """ITC global attention: the first ``num_global_blocks`` block rows and columns attend / are attended everywhere."""
        n = layout.shape[1]
        if n < self.num_global_blocks:
            raise ValueError(f"Number of global blocks, {self.num_global_blocks}, must be smaller than overall number of "
                             f"blocks in a row, {n}!")
        r, c = _grid(n)
        g = self.num_global_blocks
        layout[h] |= ((c < g) | (r < g)).expand(n, n).long()
        if self.attention == "unidirectional":
            layout[h] &= (c <= r).long()
        return layout


class BSLongformerSparsityConfig(SparsityConfig):

    def __init__(self, num_heads, block=16, different_layout_per_head=False, num_sliding_window_blocks=3,
                 global_block_indices=[0], global_block_end_indices=None, attention="bidirectional"):
        super().__init__(num_heads, block, different_layout_per_head)
        self.num_sliding_window_blocks = num_sliding_window_blocks
        self.global_block_indices = global_block_indices
        self.attention = attention
        if global_block_end_indices is not None:
            if len(global_block_indices) != len(global_block_end_indices):
                raise ValueError("global block start/end index lists must have the same length")
            for s, e in zip(global_block_indices, global_block_end_indices):
                if s >= e:
                    raise ValueError(f"Global block start index, {s}, must be smaller than global block end index, {e}!")
        self.global_block_end_indices = global_block_end_indices

    def make_layout(self, seq_len):
        layout = self.setup_layout(seq_len)
        n = layout.shape[1]
        if n < self.num_sliding_window_blocks:
            raise ValueError(f"Number of sliding window blocks, {self.num_sliding_window_blocks}, must be smaller than "
                             f"overall number of blocks in a row, {n}!")
        r, c = _grid(n)
        w = self.num_sliding_window_blocks // 2
        ends = self.global_block_end_indices or [i + 1 for i in self.global_block_indices]
        for h in range(self.num_layout_heads):
            m = (c >= r - w) & (c <= r + w)
            for s, e in zip(self.global_block_indices, ends):
                if s < n:
                    e = min(e, n)
                    m |= ((c >= s) & (c < e)).expand(n, n)
                    m |= ((r >= s) & (r < e)).expand(n, n)
            if self.attention == "unidirectional":
                m &= c <= r
            layout[h] = m.long()
        return self.check_and_propagate_first_head_layout(layout)

    def set_sliding_window_layout(self, h, layout):
        """OR the band ``|row - col| <= num_sliding_window_blocks // 2`` into head ``h``."""
        n = layout.shape[1]
        if n < self.num_sliding_window_blocks:
            raise ValueError(f"Number of sliding window blocks, {self.num_sliding_window_blocks}, must be smaller than "
                             f"overall number of blocks in a row, {n}!")
        r, c = _grid(n)
        w = self.num_sliding_window_blocks // 2
        layout[h] |= ((c >= r - w) & (c <= r + w)).long()
        return layout

    def set_global_layout(self, h, layout):
        """OR the global block rows and columns ``[start, end)`` into head ``h``."""
        n = layout.shape[1]
        r, c = _grid(n)
        ends = self.global_block_end_indices or [i + 1 for i in self.global_block_indices]
        for s0, e in zip(self.global_block_indices, ends):
            if s0 < n:
                e = min(e, n)
                layout[h] |= (((c >= s0) & (c < e)) | ((r >= s0) & (r < e))).expand(n, n).long()
        if self.attention == "unidirectional":
            layout[h] &= (c <= r).long()
        return layout


class LocalSlidingWindowSparsityConfig(SparsityConfig):

    def __init__(self, num_heads, block=16, num_sliding_window_blocks=3, attention="unidirectional"):
        super().__init__(num_heads, block)
        self.num_sliding_window_blocks = num_sliding_window_blocks
        self.attention = attention

    def make_layout(self, seq_len):
        layout = self.setup_layout(seq_len)
        n = layout.shape[1]
        if n < self.num_sliding_window_blocks:
            raise ValueError(f"Number of sliding window blocks, {self.num_sliding_window_blocks}, must be smaller than "
                             f"overall number of blocks in a row, {n}!")
        r, c = _grid(n)
        w = self.num_sliding_window_blocks // 2
        hi = r if self.attention == "unidirectional" else r + w
        layout[:] = ((c >= r - w) & (c <= hi)).long()
        return layout
