"""Linear layers that cooperate with the ZeRO flat gradient buffers.

* :func:`flat_linear` -- ``y = x W^T (+ b)``; its backward computes ``dW`` with an ``out=`` GEMM directly
  into the unit's flat gradient buffer (``ZeroShardedOptimizer.grad_view_for``) and never saves the
  gathered weight (it is re-read from the parameter at backward time, so ZeRO-3 can release it after
  forward).  Role parity: reference ``runtime/zero/linear.py:50 LinearFunctionForZeroStage3`` (P17)
  plus the bucket copy in ``stage3.py:1275 __add_grad_to_ipg_bucket`` which this design removes.
* :func:`chunked_linear_xent` -- LM head + softmax cross-entropy over token chunks; logits gradient is
  produced in place by the fused kernel, ``dh`` and ``dW`` are accumulated per chunk, so the
  ``[tokens, vocab]`` logits tensor is never materialised (role: reference ``FPDT_LogitsLoss``).
"""
import torch
import torch.nn.functional as F

from deepspeed_b200.ops.kernels.transformer_ops import softmax_xent_fwd_bwd


def _zo_of(w):
    ref = getattr(w, "_ds_zero", None)
    zo = ref() if ref is not None else None
    if zo is None or getattr(w, "ds_no_direct_grad", False) or not w.requires_grad:
        return None
    return zo


def _gemm(a, b_t):
    """a [M, K] @ b_t[N, K]^T -> [M, N] through the active GEMM backend."""
    from deepspeed_b200.ops import gemm
    return gemm.matmul_nt(a, b_t)


class _FlatLinearFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.weight_ref = weight
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x)
        x2 = x.reshape(-1, x.shape[-1])
        y = _gemm(x2, weight)
        if bias is not None:
            y = y + bias
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        (x, ) = ctx.saved_tensors
        w = ctx.weight_ref
        dy2 = dy.reshape(-1, dy.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        dx = None
        if ctx.needs_input_grad[0]:
            from deepspeed_b200.ops import gemm
            dx = gemm.matmul_nn(dy2, w).view(x.shape)
        dw = None
        if ctx.needs_input_grad[1]:
            zo = _zo_of(w)
            if zo is not None:
                gv = zo.grad_view_for(w)
                if gv.dtype == dy2.dtype:
                    if zo.grad_is_fresh(w):
                        from deepspeed_b200.ops import gemm
                        gemm.matmul_tn(dy2, x2, out=gv)
                    else:
                        gv.addmm_(dy2.t(), x2)
                else:
                    g = torch.mm(dy2.t(), x2)
                    gv.copy_(g) if zo.grad_is_fresh(w) else gv.add_(g)
                zo.mark_grad_ready(w)
            else:
                dw = torch.mm(dy2.t(), x2)
        db = dy2.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return dx, dw, db


def flat_linear(x, weight, bias=None):
    if not torch.is_grad_enabled() or not (x.requires_grad or weight.requires_grad):
        x2 = x.reshape(-1, x.shape[-1])
        y = _gemm(x2, weight)
        if bias is not None:
            y = y + bias
        return y.view(*x.shape[:-1], weight.shape[0])
    return _FlatLinearFn.apply(x, weight, bias)


class _ChunkedLinearXent(torch.autograd.Function):

    @staticmethod
    def forward(ctx, h, weight, labels, chunk, ignore_index, assumed_scale):
        T, H = h.shape
        valid = (labels != ignore_index)
        n_valid = valid.sum().clamp(min=1).to(torch.float32)
        inv_n = (assumed_scale / n_valid).reshape(1)  # device scalar: no host sync
        need_dh = h.requires_grad
        need_dw = weight.requires_grad
        dh = torch.empty_like(h) if need_dh else None
        dw_tmp = None
        if need_dw:
            # per-call scratch (the caching allocator recycles the block): several forwards may precede
            # their backwards, so this must not be shared between calls
            dw_tmp = torch.empty(weight.shape, dtype=weight.dtype, device=weight.device)
        total = torch.zeros((), dtype=torch.float32, device=h.device)
        first = True
        for s in range(0, T, chunk):
            e = min(s + chunk, T)
            hc = h[s:e]
            logits = _gemm(hc, weight)  # [c, V]
            loss_rows, grad = softmax_xent_fwd_bwd(logits, labels[s:e].contiguous(), 1.0, inv_n, ignore_index, True)
            total = total + loss_rows.sum()
            if need_dh:
                from deepspeed_b200.ops import gemm
                gemm.matmul_nn(grad, weight, out=dh[s:e])
            if need_dw:
                if first:
                    from deepspeed_b200.ops import gemm
                    gemm.matmul_tn(grad, hc, out=dw_tmp)
                else:
                    dw_tmp.addmm_(grad.t(), hc)
            first = False
        ctx.weight_ref = weight
        ctx.assumed_scale = assumed_scale
        ctx.need_dw = need_dw
        ctx.dh = dh
        ctx.dw_tmp = dw_tmp
        return total / n_valid

    @staticmethod
    def backward(ctx, dloss):
        w = ctx.weight_ref
        r = (dloss / ctx.assumed_scale).to(torch.float32)  # == 1 when the engine's multiplier was assumed
        dh = None
        if ctx.dh is not None:
            dh = ctx.dh.mul_(r.to(ctx.dh.dtype))
        dw = None
        if ctx.need_dw:
            zo = _zo_of(w)
            if zo is not None:
                gv = zo.grad_view_for(w)
                if zo.grad_is_fresh(w):
                    torch.mul(ctx.dw_tmp, r.to(ctx.dw_tmp.dtype), out=gv) if gv.dtype == ctx.dw_tmp.dtype else gv.copy_(
                        ctx.dw_tmp.float() * r)
                else:
                    gv.add_((ctx.dw_tmp.float() * r).to(gv.dtype))
                zo.mark_grad_ready(w)
            else:
                dw = ctx.dw_tmp * r.to(ctx.dw_tmp.dtype)
        ctx.dh = None
        return dh, dw, None, None, None, None


def chunked_linear_xent(h, weight, labels, chunk=2048, ignore_index=-100, assumed_scale=1.0):
    """Mean cross entropy of ``softmax(h W^T)`` vs ``labels`` without materialising the logits."""
    if not torch.is_grad_enabled():
        total = torch.zeros((), dtype=torch.float32, device=h.device)
        n = (labels != ignore_index).sum().clamp(min=1)
        for s in range(0, h.shape[0], chunk):
            e = min(s + chunk, h.shape[0])
            logits = _gemm(h[s:e], weight)
            rows, _ = softmax_xent_fwd_bwd(logits, labels[s:e].contiguous(), 1.0, None, ignore_index, False)
            total = total + rows.sum()
        return total / n
    return _ChunkedLinearXent.apply(h, weight, labels, chunk, ignore_index, float(assumed_scale) or 1.0)
