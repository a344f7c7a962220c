"""Python face of ``csrc/cuda/attn_sm100.cu`` -- the tcgen05 / TMEM training attention (head dim 128, bf16, GQA, causal).

``fwd`` / ``bwd`` work on 2-D ``[B*S, heads*128]`` views (row stride free), so the packed QKV projection output is consumed
in place: Q, K and V are column ranges of ONE buffer and the backward writes dQ | dK | dV into one packed buffer that the
QKV projection's backward GEMMs read directly (no transposes, no concatenation).
"""
import math

import torch

from deepspeed_b200.ops import native as N

HEAD_DIM = 128
BLOCK = 128


def supports(qkv2d, hq, hkv, d, S) -> bool:
    return (qkv2d.is_cuda and qkv2d.dtype == torch.bfloat16 and d == HEAD_DIM and S % BLOCK == 0 and hq % hkv == 0
            and qkv2d.stride(-1) == 1 and qkv2d.stride(0) % 8 == 0 and qkv2d.data_ptr() % 16 == 0)


def supports_fwd(x2d, hq, hkv, d, B, S) -> bool:
    """Forward-only (inference prefill): any sequence length when one sequence is launched at a time."""
    return (x2d.is_cuda and x2d.dtype == torch.bfloat16 and d == HEAD_DIM and hq % hkv == 0 and x2d.stride(-1) == 1
            and x2d.stride(0) % 8 == 0 and x2d.data_ptr() % 16 == 0 and (B == 1 or S % BLOCK == 0))


def fwd(q, k, v, B, S, hq, hkv, causal=True, scale=None, out=None, need_lse=True):
    """``q [B*S, hq*128]``, ``k / v [B*S, hkv*128]`` (2-D views, any row stride) -> ``(o [B*S, hq*128], lse [B, hq, S])``.
    ``S`` may be any length when ``B == 1`` (a ragged last block is masked inside the kernel)."""
    T = B * S
    if out is None:
        out = torch.empty(T, hq * HEAD_DIM, dtype=torch.bfloat16, device=q.device)
    lse = torch.empty(B, hq, S, dtype=torch.float32, device=q.device) if need_lse else None
    scale = float(scale) if scale is not None else 1.0 / math.sqrt(HEAD_DIM)
    rc = N.cuda().dsb_attn_fwd_bf16(N.ptr(q), N.ptr(k), N.ptr(v), N.ptr(out), N.ptr(lse), B, S, hq, hkv, HEAD_DIM, q.stride(0),
                                    k.stride(0), v.stride(0), out.stride(0), q.shape[1], k.shape[1], v.shape[1],
                                    N.c_f(scale), int(bool(causal)), N.stream())
    N.check(rc, "attn_fwd_bf16")
    return out, lse


def split_packed(qkv2d, hq, hkv, d=HEAD_DIM):
    """Column views (no copies) of a packed ``[T, (hq + 2 hkv) * d]`` projection output."""
    return qkv2d[:, :hq * d], qkv2d[:, hq * d:(hq + hkv) * d], qkv2d[:, (hq + hkv) * d:]


def bwd(d_o, q, k, v, o, lse, B, S, hq, hkv, causal=True, scale=None, dq=None, dk=None, dv=None):
    """Gradients of :func:`fwd`.  ``dq / dk / dv`` may be column views of one packed ``[T, (hq + 2 hkv) * 128]`` buffer."""
    T = B * S
    dev = q.device
    if dq is None:
        dq = torch.empty(T, hq * HEAD_DIM, dtype=torch.bfloat16, device=dev)
    if dk is None:
        dk = torch.empty(T, hkv * HEAD_DIM, dtype=torch.bfloat16, device=dev)
    if dv is None:
        dv = torch.empty(T, hkv * HEAD_DIM, dtype=torch.bfloat16, device=dev)
    if d_o.stride(-1) != 1:
        d_o = d_o.contiguous()
    work = torch.empty(2 * B * hq * S, dtype=torch.float32, device=dev)
    scale = float(scale) if scale is not None else 1.0 / math.sqrt(HEAD_DIM)
    rc = N.cuda().dsb_attn_bwd_bf16(N.ptr(q), N.ptr(k), N.ptr(v), N.ptr(o), N.ptr(d_o), N.ptr(lse), N.ptr(dq), N.ptr(dk),
                                    N.ptr(dv), N.ptr(work), B, S, hq, hkv, HEAD_DIM, q.stride(0), k.stride(0), v.stride(0),
                                    o.stride(0), d_o.stride(0), dq.stride(0), dk.stride(0), dv.stride(0), q.shape[1],
                                    k.shape[1], v.shape[1], N.c_f(scale), int(bool(causal)), N.stream())
    N.check(rc, "attn_bwd_bf16")
    return dq, dk, dv


class _PackedAttention(torch.autograd.Function):
    """Causal GQA attention on the packed QKV projection output (RoPE applied in place first): the backward leaves
    dQ | dK | dV in ONE packed buffer -- exactly the layout the QKV projection's backward GEMMs consume."""

    @staticmethod
    def forward(ctx, qkv, B, S, hq, hkv, rope, positions):
        from deepspeed_b200.ops.kernels.transformer_ops import rope_qk_inplace
        if rope is not None:
            rope_qk_inplace(qkv, hq, hkv, HEAD_DIM, rope, positions, S, backward=False)
        q, k, v = split_packed(qkv, hq, hkv)
        o, lse = fwd(q, k, v, B, S, hq, hkv, causal=True)
        ctx.save_for_backward(qkv, o, lse)
        ctx.meta = (B, S, hq, hkv, rope, positions)
        return o

    @staticmethod
    def backward(ctx, d_o):
        from deepspeed_b200.ops.kernels.transformer_ops import rope_qk_inplace
        qkv, o, lse = ctx.saved_tensors
        B, S, hq, hkv, rope, positions = ctx.meta
        q, k, v = split_packed(qkv, hq, hkv)
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = split_packed(dqkv, hq, hkv)
        bwd(d_o.reshape(B * S, hq * HEAD_DIM), q, k, v, o, lse, B, S, hq, hkv, causal=True, dq=dq, dk=dk, dv=dv)
        if rope is not None:
            rope_qk_inplace(dqkv, hq, hkv, HEAD_DIM, rope, positions, S, backward=True)
        return dqkv, None, None, None, None, None, None


def packed_causal_attention(qkv2d, B, S, hq, hkv, d, rope=None, positions=None):
    assert d == HEAD_DIM
    return _PackedAttention.apply(qkv2d, B, S, hq, hkv, rope, positions)
