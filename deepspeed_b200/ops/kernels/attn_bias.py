"""Python face of ``csrc/cuda/attn_bias.cu``: flash attention with two additive biases and an optional block-sparsity
layout for head dims 16 / 32 / 64 (bf16 / fp16), forward and backward including both bias gradients.

Operands are 4-D *logical* ``[NB, H, L, D]`` tensors with free batch / head / row strides (last dim contiguous), so the
Evoformer ``[*, L, H, D]`` layout and the sparse-attention ``[B, H, L, D]`` layout are both consumed in place.

* ``bias1`` ``[NB or 1, Lk]``   -- per-key additive bias (Evoformer mask bias, key-padding mask)
* ``bias2`` ``[B2, H or 1, Lq, Lk]`` with ``B2 * b2_div == NB`` -- pair bias shared by ``b2_div`` consecutive batch entries
* ``layout`` ``[H or 1, Lq / block, Lk / block]`` 0/1 -- score blocks that exist (others are -inf and skipped tile-wise)
"""
import ctypes
import math

import torch

from deepspeed_b200.ops import native as N

HEAD_DIMS = (16, 32, 64)
MAX_GRID = 65535


def supported(q, k=None, v=None) -> bool:
    return (q.is_cuda and q.dtype in (torch.bfloat16, torch.float16) and q.dim() == 4 and q.shape[-1] in HEAD_DIMS
            and q.shape[0] <= MAX_GRID and q.shape[1] <= MAX_GRID and (k is None or k.shape[-1] == q.shape[-1])
            and (v is None or v.shape[-1] == q.shape[-1]))


def _ok_strides(t):
    return t.stride(-1) == 1 and all(s % 8 == 0 for s in t.stride()[:-1]) and t.data_ptr() % 16 == 0


def _prep(t):
    return t if _ok_strides(t) else t.contiguous()


def _call(q, k, v, o, lse, bias1, bias2, b2_div, layout, block, causal, scale, d_o=None, dq=None, dk=None, dv=None, delta=None,
          db1=None, db2=None):
    NB, H, Lq, D = q.shape
    Lk = k.shape[2]
    lay_nq = lay_nk = 0
    lay_h = 0
    if layout is not None:
        lay_nq, lay_nk = layout.shape[-2:]
        lay_h = lay_nq * lay_nk if layout.shape[0] > 1 else 0
    b1_b = 0
    if bias1 is not None:
        b1_b = Lk if bias1.shape[0] > 1 else 0
    b2 = [0, 0, 0]
    g2 = [0, 0]
    if bias2 is not None:
        b2 = [bias2.stride(0) if bias2.shape[0] > 1 else 0, bias2.stride(1) if bias2.shape[1] > 1 else 0, bias2.stride(2)]
        g2 = [bias2.shape[1] * Lq * Lk, Lq * Lk if bias2.shape[1] > 1 else 0]
    ip = (ctypes.c_int32 * 12)(NB, H, Lq, Lk, D, N.dt(q), int(b2_div), int(block or 0), lay_nq, lay_nk, int(bool(causal)),
                               int(d_o is not None))
    st = (ctypes.c_int64 * 19)(q.stride(0), q.stride(1), q.stride(2), k.stride(0), k.stride(1), k.stride(2), v.stride(0),
                               v.stride(1), v.stride(2), o.stride(0), o.stride(1), o.stride(2), b1_b, b2[0], b2[1], b2[2], g2[0],
                               g2[1], lay_h)
    rc = N.cuda().dsb_attn_bias(N.ptr(q), N.ptr(k), N.ptr(v), N.ptr(o), N.ptr(lse), N.ptr(bias1), N.ptr(bias2), N.ptr(layout),
                                N.ptr(d_o), N.ptr(dq), N.ptr(dk), N.ptr(dv), N.ptr(delta), N.ptr(db1), N.ptr(db2), ip, st,
                                ctypes.c_float(scale), N.stream())
    N.check(rc, "attn_bias")


def _norm_args(q, k, v, bias1, bias2, layout):
    q, k, v = _prep(q), _prep(k), _prep(v)
    if bias1 is not None:
        bias1 = bias1.to(q.dtype).contiguous()
        assert bias1.dim() == 2 and bias1.shape[0] in (1, q.shape[0]) and bias1.shape[1] == k.shape[2], bias1.shape
    b2_div = 1
    if bias2 is not None:
        bias2 = bias2.to(q.dtype)
        if bias2.stride(-1) != 1:
            bias2 = bias2.contiguous()
        assert bias2.dim() == 4 and bias2.shape[1] in (1, q.shape[1]) and bias2.shape[2:] == (q.shape[2], k.shape[2]), bias2.shape
        assert q.shape[0] % bias2.shape[0] == 0, "bias2 batch must divide the attention batch"
        b2_div = q.shape[0] // bias2.shape[0]
    if layout is not None:
        layout = layout.to(device=q.device, dtype=torch.uint8).contiguous()
        if layout.dim() == 2:
            layout = layout[None]
        assert layout.shape[0] in (1, q.shape[1]), layout.shape
    return q, k, v, bias1, bias2, b2_div, layout


def forward(q, k, v, bias1=None, bias2=None, layout=None, block=0, causal=False, scale=None, need_lse=True):
    """-> ``(o [NB, H, Lq, D] with q's strides, lse [NB, H, Lq] fp32)``."""
    q, k, v, bias1, bias2, b2_div, layout = _norm_args(q, k, v, bias1, bias2, layout)
    scale = float(scale) if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    o = torch.empty_like(q)
    if not _ok_strides(o) or o.stride() != q.stride():
        o = torch.empty(q.shape, dtype=q.dtype, device=q.device)
    lse = torch.empty(q.shape[:3], dtype=torch.float32, device=q.device) if need_lse else None
    _call(q, k, v, o, lse, bias1, bias2, b2_div, layout, block, causal, scale)
    return o, lse


def backward(d_o, q, k, v, o, lse, bias1=None, bias2=None, layout=None, block=0, causal=False, scale=None, need_db1=False,
             need_db2=False):
    """-> ``(dq, dk, dv, db1 | None, db2 | None)``; bias gradients are fp32 sums cast to the bias dtype."""
    q, k, v, bias1, bias2, b2_div, layout = _norm_args(q, k, v, bias1, bias2, layout)
    scale = float(scale) if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    if d_o.stride() != o.stride() or not _ok_strides(d_o):
        d_o = torch.empty_like(o).copy_(d_o)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    for a, b in ((dq, q), (dk, k), (dv, v)):
        assert a.stride() == b.stride(), "gradient buffers must share the operand strides"
    delta = torch.empty_like(lse)
    db1 = torch.zeros(bias1.shape, dtype=torch.float32, device=q.device) if (need_db1 and bias1 is not None) else None
    db2 = torch.zeros(bias2.shape, dtype=torch.float32, device=q.device) if (need_db2 and bias2 is not None) else None
    _call(q, k, v, o, lse, bias1, bias2, b2_div, layout, block, causal, scale, d_o, dq, dk, dv, delta, db1, db2)
    return dq, dk, dv, db1, db2


class BiasedAttention(torch.autograd.Function):

    @staticmethod
    def forward(ctx, q, k, v, bias1, bias2, layout, block, causal, scale):
        o, lse = forward(q, k, v, bias1, bias2, layout, block, causal, scale)
        ctx.save_for_backward(q, k, v, o, lse, bias1, bias2, layout)
        ctx.cfg = (block, causal, scale)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse, bias1, bias2, layout = ctx.saved_tensors
        block, causal, scale = ctx.cfg
        dq, dk, dv, db1, db2 = backward(d_o, q, k, v, o, lse, bias1, bias2, layout, block, causal, scale,
                                        need_db1=bias1 is not None and ctx.needs_input_grad[3],
                                        need_db2=bias2 is not None and ctx.needs_input_grad[4])
        if db1 is not None:
            db1 = db1.to(bias1.dtype)
        if db2 is not None:
            db2 = db2.to(bias2.dtype)
        return dq, dk, dv, db1, db2, None, None, None, None


def biased_attention(q, k, v, bias1=None, bias2=None, layout=None, block=0, causal=False, scale=None):
    """Differentiable entry point (q / k / v and both biases)."""
    return BiasedAttention.apply(q, k, v, bias1, bias2, layout, block, causal, scale)
