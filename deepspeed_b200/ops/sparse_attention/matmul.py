"""Block-sparse matmul (reference Triton ``matmul.py``): modes ``sdd`` (dense x dense -> sparse blocks),
``dsd`` (sparse x dense -> dense), ``dds``.  Sparse operands are ``[B, nnz_blocks, block, block]`` in the row-major
order of ``layout.nonzero()``.  Implemented with batched tensor-core GEMMs over gathered blocks."""
import torch


class MatMul:

    def __init__(self, layout, block, mode, trans_a=False, trans_b=False, bench=False):
        if mode not in ("sdd", "dsd", "dds"):
            raise NotImplementedError("Supported modes are: sdd, dsd, dds")
        self.layout, self.block, self.mode, self.trans_a, self.trans_b = layout, block, mode, trans_a, trans_b
        self.idx = layout.nonzero()  # [nnz, 3] (head, row, col)
        self.spdims = layout.shape

    def __call__(self, a, b):
        blk = self.block
        h, r, c = (self.idx[:, i].to(a.device) for i in range(3))
        if self.mode == "sdd":
            a_ = a.transpose(-1, -2) if self.trans_a else a
            b_ = b.transpose(-1, -2) if self.trans_b else b
            B, H, M, K = a_.shape
            ab = a_.reshape(B, H, M // blk, blk, K)[:, h, r]            # [B, nnz, blk, K]
            bb = b_.reshape(B, H, K, b_.shape[-1] // blk, blk).permute(0, 1, 3, 2, 4)[:, h, c]  # [B, nnz, K, blk]
            return torch.matmul(ab, bb)
        if self.mode == "dsd":
            # a sparse [B, nnz, blk, blk], b dense [B, H, K, N]
            b_ = b.transpose(-1, -2) if self.trans_b else b
            B, H, K, N = b_.shape
            a_ = a.transpose(-1, -2) if self.trans_a else a
            rr, cc = (c, r) if self.trans_a else (r, c)
            bb = b_.reshape(B, H, K // blk, blk, N)[:, h, cc]           # [B, nnz, blk, N]
            prod = torch.matmul(a_, bb)
            nrow = self.spdims[1 if not self.trans_a else 2]
            out = torch.zeros(B, H * nrow, blk, N, dtype=prod.dtype, device=prod.device)
            out.index_add_(1, h * nrow + rr, prod)
            return out.view(B, H, nrow * blk, N)
        # dds: a dense [B,H,M,K], b sparse
        a_ = a.transpose(-1, -2) if self.trans_a else a
        B, H, M, K = a_.shape
        b_ = b.transpose(-1, -2) if self.trans_b else b
        rr, cc = (c, r) if self.trans_b else (r, c)
        ab = a_.reshape(B, H, M, K // blk, blk).permute(0, 1, 3, 2, 4)[:, h, rr]  # [B, nnz, M, blk]
        prod = torch.matmul(ab, b_)
        ncol = self.spdims[2 if not self.trans_b else 1]
        out = torch.zeros(B, H * ncol, M, blk, dtype=prod.dtype, device=prod.device)
        out.index_add_(1, h * ncol + cc, prod)
        return out.view(B, H, ncol, M, blk).permute(0, 1, 3, 2, 4).reshape(B, H, M, ncol * blk)
