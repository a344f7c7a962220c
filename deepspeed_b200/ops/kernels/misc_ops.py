"""Python face of ``csrc/cuda/misc.cu``: dropout (+bias +residual), masked attention softmax, QKV bias+permute,
random-LTD token sort / gather / scatter / mask slicing, NHWC bias-add.  Host tensors take equivalent torch
code.  Reference counterparts: N7 ``dropout_kernels.cu`` / ``softmax_kernels.cu`` / ``transform_kernels.cu``,
N13 ``csrc/spatial``, N14 ``csrc/random_ltd``."""
import ctypes

import torch

from deepspeed_b200.ops import native as N


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_seed_state = {"offset": 0}


def _next_offset(n):
    o = _seed_state["offset"]
    _seed_state["offset"] = o + n
    return o


class _Dropout(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, bias, residual, p, seed):
        n = x.numel()
        cols = x.shape[-1]
        if seed is None:
            seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFF
        off = _next_offset(n)
        ctx.p, ctx.seed, ctx.off, ctx.has_bias, ctx.has_res = p, seed, off, bias is not None, residual is not None
        if x.is_cuda:
            xc = x.contiguous()
            y = torch.empty_like(xc)
            rc = N.cuda().dsb_dropout(_p(xc), _p(bias), _p(residual.contiguous() if residual is not None else None), _p(y),
                                      ctypes.c_void_p(0), ctypes.c_int64(n), cols, ctypes.c_float(p),
                                      ctypes.c_uint64(seed), ctypes.c_uint64(off), N.dt(x), N.stream())
            N.check(rc, "dropout")
            ctx.mask = None
            return y
        g = torch.Generator().manual_seed((seed + off) & 0x7FFFFFFF)
        keep = (torch.rand(x.shape, generator=g) >= p)
        ctx.mask = keep
        f = x.float() + (bias.float() if bias is not None else 0)
        f = f * keep / (1.0 - p)
        if residual is not None:
            f = f + residual.float()
        return f.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        if dy.is_cuda:
            dyc = dy.contiguous()
            dx = torch.empty_like(dyc)
            rc = N.cuda().dsb_dropout_bwd(_p(dyc), ctypes.c_void_p(0), _p(dx), ctypes.c_int64(dy.numel()),
                                          ctypes.c_float(ctx.p), ctypes.c_uint64(ctx.seed), ctypes.c_uint64(ctx.off),
                                          N.dt(dy), N.stream())
            N.check(rc, "dropout_bwd")
        else:
            dx = (dy.float() * ctx.mask / (1.0 - ctx.p)).to(dy.dtype)
        db = dx.reshape(-1, dx.shape[-1]).sum(0) if ctx.has_bias else None
        return dx, db, (dy if ctx.has_res else None), None, None


def dropout(x, p, training=True, bias=None, residual=None, seed=None):
    """``dropout(x + bias) + residual``; the keep mask is regenerated (not stored) in backward."""
    if not training or p == 0.0:
        y = x if bias is None else x + bias
        return y if residual is None else y + residual
    return _Dropout.apply(x, bias, residual, float(p), seed)


class _AttnSoftmax(torch.autograd.Function):

    @staticmethod
    def forward(ctx, scores, mask, alibi, scale, causal, window):
        b, h, sq, sk = scores.shape
        ctx.scale = scale
        if scores.is_cuda:
            s = scores.contiguous().clone()
            mask_sq = mask.shape[-2] if mask is not None else 1
            rc = N.cuda().dsb_attn_softmax(_p(s), _p(mask.contiguous() if mask is not None else None),
                                           _p(alibi.float().contiguous() if alibi is not None else None), b, h, sq, sk,
                                           ctypes.c_float(scale), int(causal), int(window), mask_sq, N.dt(s), N.stream())
            N.check(rc, "attn_softmax")
        else:
            v = scores.float() * scale
            if mask is not None:
                v = v + mask.float()
            i_abs = torch.arange(sq)[:, None] + (sk - sq)
            j = torch.arange(sk)[None, :]
            if alibi is not None:
                v = v + alibi.float().view(1, h, 1, 1) * (j - i_abs).float()
            dead = torch.zeros(sq, sk, dtype=torch.bool)
            if causal:
                dead |= j > i_abs
            if window > 0:
                dead |= j <= i_abs - window
            v = v.masked_fill(dead, float("-inf"))
            s = torch.softmax(v, -1).nan_to_num(0.0).to(scores.dtype)
        ctx.save_for_backward(s)
        return s

    @staticmethod
    def backward(ctx, dp):
        (p, ) = ctx.saved_tensors
        if dp.is_cuda:
            d = dp.contiguous().clone()
            rows = p.numel() // p.shape[-1]
            rc = N.cuda().dsb_attn_softmax_bwd(_p(d), _p(p), ctypes.c_int64(rows), p.shape[-1], ctypes.c_float(ctx.scale),
                                               N.dt(d), N.stream())
            N.check(rc, "attn_softmax_bwd")
        else:
            pf, df = p.float(), dp.float()
            d = ((df - (df * pf).sum(-1, keepdim=True)) * pf * ctx.scale).to(dp.dtype)
        return d, None, None, None, None, None


def attn_softmax(scores, mask=None, alibi=None, scale=1.0, causal=False, window=0):
    return _AttnSoftmax.apply(scores, mask, alibi, float(scale), bool(causal), int(window))


def bias_transform_0213(x, bias, b, s, n3, h, d):
    """[b, s, n3, h, d] (+bias) -> [n3, b, h, s, d]"""
    if x.is_cuda:
        out = torch.empty(n3, b, h, s, d, dtype=x.dtype, device=x.device)
        rc = N.cuda().dsb_bias_transform_0213(_p(x.contiguous()), _p(bias), _p(out), b, s, n3, h, d, N.dt(x), N.stream())
        N.check(rc, "bias_transform_0213")
        return out
    v = x.view(b, s, n3, h, d)
    if bias is not None:
        v = v + bias.view(n3, h, d)
    return v.permute(2, 0, 3, 1, 4).contiguous()


def transform4d_0213(x):
    """[b, h, s, d] -> [b, s, h, d]"""
    b, h, s, d = x.shape
    if x.is_cuda:
        out = torch.empty(b, s, h, d, dtype=x.dtype, device=x.device)
        rc = N.cuda().dsb_transform4d_0213(_p(x.contiguous()), _p(out), b, h, s, d, N.dt(x), N.stream())
        N.check(rc, "transform4d_0213")
        return out
    return x.permute(0, 2, 1, 3).contiguous()


def token_sort_(idx):
    """Sort each row of int32 ``idx`` [rows, k] ascending, in place."""
    rows, k = idx.shape
    if idx.is_cuda:
        rc = N.cuda().dsb_token_sort(_p(idx), rows, k, N.stream())
        N.check(rc, "token_sort")
        return idx
    idx.copy_(idx.sort(dim=-1).values)
    return idx


def token_gather(x, idx):
    """x [B, S, H], idx [B, k] -> [B, k, H]"""
    B, S, H = x.shape
    k = idx.shape[1]
    if x.is_cuda:
        out = torch.empty(B, k, H, dtype=x.dtype, device=x.device)
        rc = N.cuda().dsb_token_gather(_p(x.contiguous()), _p(idx), _p(out), B, S, k, H, 0, N.dt(x), N.stream())
        N.check(rc, "token_gather")
        return out
    return torch.gather(x, 1, idx.long()[..., None].expand(B, k, H))


def token_scatter_(full, part, idx):
    """full[b, idx[b, j]] = part[b, j] in place; returns ``full``."""
    B, S, H = full.shape
    k = idx.shape[1]
    if full.is_cuda:
        rc = N.cuda().dsb_token_gather(_p(part.contiguous()), _p(idx), _p(full), B, S, k, H, 1, N.dt(full), N.stream())
        N.check(rc, "token_scatter")
        return full
    full.scatter_(1, idx.long()[..., None].expand(B, k, H), part)
    return full


def mask_gather(mask, idx):
    """mask [B, 1, S, S], idx [B, k] -> [B, 1, k, k]"""
    B, _, S, _ = mask.shape
    k = idx.shape[1]
    if mask.is_cuda:
        out = torch.empty(B, 1, k, k, dtype=mask.dtype, device=mask.device)
        rc = N.cuda().dsb_mask_gather(_p(mask.contiguous()), _p(idx), _p(out), B, S, k, N.dt(mask), N.stream())
        N.check(rc, "mask_gather")
        return out
    i = idx.long()
    return mask[torch.arange(B)[:, None, None], 0, i[:, :, None], i[:, None, :]].unsqueeze(1)


def nhwc_bias_add(x, bias, other=None, other_bias=None):
    C = x.shape[-1]
    if x.is_cuda and x.dtype in (torch.float16, torch.bfloat16, torch.float32) and x.numel() % 8 == 0 and C % 8 == 0:
        xc = x.contiguous()
        y = torch.empty_like(xc)
        rc = N.cuda().dsb_nhwc_bias_add(_p(xc), _p(bias), _p(other.contiguous() if other is not None else None),
                                        _p(other_bias), _p(y), ctypes.c_int64(x.numel()), C, N.dt(x), N.stream())
        N.check(rc, "nhwc_bias_add")
        return y
    y = x + bias
    if other is not None:
        y = y + (other + other_bias if other_bias is not None else other)
    return y
