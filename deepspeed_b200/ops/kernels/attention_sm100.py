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


def fwd(q, k, v, B, S, hq, hkv, causal=True, scale=None, out=None, need_lse=True):
    """``q [B*S, hq*128]``, ``k / v [B*S, hkv*128]`` (2-D views, any row stride) -> ``(o [B*S, hq*128], lse [B, hq, S])``."""
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
