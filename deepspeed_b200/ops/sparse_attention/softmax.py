"""Block-sparse softmax over ``[B, nnz, block, block]`` scores (reference Triton ``softmax.py``): rows are
normalised across all blocks of the same (head, block-row)."""
import torch


class Softmax:

    def __init__(self, layout, block, bench=False):
        self.layout, self.block = layout, block
        self.idx = layout.nonzero()
        n_rows = layout.shape[1]
        self.row_id = self.idx[:, 0] * n_rows + self.idx[:, 1]
        self.n_groups = layout.shape[0] * n_rows

    def __call__(self, x, scale=1.0, rpe=None, key_padding_mask=None, attn_mask=None, key_padding_mask_mode="add",
                 attn_mask_mode="add"):
        B, nnz, blk, _ = x.shape
        dev = x.device
        h, r, c = (self.idx[:, i].to(dev) for i in range(3))
        v = x.float() * scale
        if rpe is not None:
            S = rpe.shape[-1]
            rp = rpe.float().reshape(-1, S // blk, blk, S // blk, blk)
            hh = h if rp.shape[0] > 1 else torch.zeros_like(h)
            v = v + rp[hh, r, :, c][None]
        if key_padding_mask is not None:
            kp = key_padding_mask.float().view(B, -1, blk)[:, c]  # [B, nnz, blk]
            kp = kp if key_padding_mask_mode == "add" else (1.0 - kp) * -10000.0
            v = v + kp[:, :, None, :]
        if attn_mask is not None:
            S = attn_mask.shape[-1]
            am = attn_mask.float().view(S // blk, blk, S // blk, blk)[r, :, c]  # [nnz, blk, blk]
            am = am if attn_mask_mode == "add" else (1.0 - am) * -10000.0
            v = v + am[None]
        rid = self.row_id.to(dev)
        mx = torch.full((B, self.n_groups, blk), float("-inf"), device=dev)
        mx = mx.scatter_reduce(1, rid[None, :, None].expand(B, nnz, blk), v.amax(-1), reduce="amax")
        e = torch.exp(v - mx[:, rid][..., None])
        den = torch.zeros(B, self.n_groups, blk, device=dev).index_add_(1, rid, e.sum(-1))
        return (e / den[:, rid][..., None]).to(x.dtype)


def next_power_of_2(n):
    """Smallest power of two ≥ ``n`` (row-tile sizing helper of the reference softmax)."""
    n = int(n)
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def num_warps(n):
    """Warps per CTA the block-sparse softmax uses for a row of ``n`` elements: 4 / 8 / 16 by row length."""
    return 4 if n < 512 else (8 if n < 2048 else 16)
