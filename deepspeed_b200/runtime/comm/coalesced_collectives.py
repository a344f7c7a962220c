"""Batched / quantised collectives used by ZeRO (reference ``runtime/comm/coalesced_collectives.py``:
``reduce_scatter_coalesced :144``, ``all_to_all_quant_reduce :81`` (qgZ), ``all_to_all_loco_quant_reduce :30``)."""
import math
from typing import List

import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.ops.quantizer import quantizer as Q


def reduce_scatter_coalesced(tensors: List[torch.Tensor], group=None) -> List[torch.Tensor]:
    """Reduce-scatter a list of tensors with ONE collective: each tensor is padded to a multiple of the world
    size, rank r's slices are packed contiguously; returns this rank's (averaged) partition of every tensor."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    parts = [math.ceil(t.numel() / world) for t in tensors]
    total = sum(parts)
    dtype, dev = tensors[0].dtype, tensors[0].device
    packed = torch.zeros(world, total, dtype=dtype, device=dev)
    off = 0
    for t, p in zip(tensors, parts):
        flat = t.reshape(-1)
        padded = torch.zeros(p * world, dtype=dtype, device=dev)
        padded[:flat.numel()] = flat
        packed[:, off:off + p] = padded.view(world, p)
        off += p
    packed.div_(world)
    out = torch.empty(total, dtype=dtype, device=dev)
    dist.reduce_scatter_tensor(out, packed.view(-1), group=group)
    res, off = [], 0
    for t, p in zip(tensors, parts):
        lo = rank * p
        valid = max(0, min(p, t.numel() - lo))
        res.append(out[off:off + p][:valid] if valid < p else out[off:off + p])
        off += p
    return res


def _groups_for(t, group_size=2048):
    n = t.numel()
    g = max(1, n // group_size)
    while n % g:
        g -= 1
    return g


def _rank_chunks(flat, world):
    """Lay ``flat`` out as ``world`` equal chunks of whole quantisation groups (rank r's elements first in chunk r, zero
    padded): -> (padded flat, per0 = elements a rank owns, per = padded chunk length, gs = group size)."""
    n = flat.numel()
    per0 = math.ceil(n / world)
    gs = Q.aligned_group_size(per0)          # device kernels: groups of a multiple of 8 elements, never across ranks
    per = (per0 + gs - 1) // gs * gs
    if per * world != n:
        full = flat.new_zeros(per * world)
        for r in range(world):
            lo_r, hi_r = r * per0, min(n, (r + 1) * per0)
            if hi_r > lo_r:
                full[r * per:r * per + hi_r - lo_r].copy_(flat[lo_r:hi_r])
        flat = full
    return flat, per0, per, gs


def _exchange_and_reduce(q, params, world, rank, per0, per, gs, n, num_bits, dtype, group):
    """Quantised payload -> all-to-all -> dequantise -> mean over the senders: this rank's slice of the reduced tensor."""
    q_recv, p_recv = torch.empty_like(q), torch.empty_like(params)
    dist.all_to_all_single(q_recv.view(-1), q.view(-1), group=group)
    dist.all_to_all_single(p_recv.view(-1), params.view(-1), group=group)
    deq = Q.dequantize(q_recv, p_recv, (per // gs) * world, num_bits, Q.Symmetric, dtype=torch.float32)
    red = deq.view(world, per).sum(0).div_(world)
    valid = max(0, min(per0, n - rank * per0))
    return red[:valid].to(dtype)


def all_to_all_quant_reduce(tensors: List[torch.Tensor], groups: dict = None, num_bits=4) -> List[torch.Tensor]:
    """qgZ: gradients travel quantised (int4 intra-node hop, int8 inter-node hop) through all-to-alls and are
    reduced after dequantisation; on a single NVSwitch node this is one quantised all-to-all + local reduce."""
    groups = groups or {}
    local = groups.get("local")
    world = dist.get_world_size(local)
    rank = dist.get_rank(local)
    out = []
    for t in tensors:
        flat = t.reshape(-1)
        n = flat.numel()
        if world == 1:
            out.append(flat.clone())
            continue
        flat, per0, per, gs = _rank_chunks(flat, world)
        q, params = Q.quantize(flat.contiguous(), (per // gs) * world, num_bits, Q.Symmetric)
        out.append(_exchange_and_reduce(q, params, world, rank, per0, per, gs, n, num_bits, t.dtype, local))
    return out


def all_to_all_loco_quant_reduce(params, groups: dict = None, loco_param: dict = None, num_bits=4):
    """LoCo-ZeRO++: qgZ with an error-feedback buffer per tensor (``p.intra_ef_buf``) so the quantisation error is
    re-injected at the next step (reference :30).  Compensation, quantisation and the error update are ONE kernel
    (``quant.cu loco_quantize_kernel``); the buffer lives in the padded per-rank chunk layout the payload travels in."""
    loco_param = loco_param or {}
    groups = groups or {}
    local = groups.get("local")
    world, rank = dist.get_world_size(local), dist.get_rank(local)
    beta = float(loco_param.get("err_beta", 0.8))
    reset_T = int(loco_param.get("reset_T", 1024))
    outs = []
    for p in params:
        g = p.grad if hasattr(p, "grad") and p.grad is not None else p
        flat = g.reshape(-1)
        n = flat.numel()
        if world == 1:
            outs.append(flat.clone())
            continue
        flat, per0, per, gs = _rank_chunks(flat, world)
        buf = getattr(p, "intra_ef_buf", None)
        if buf is None or buf[0].numel() != flat.numel() or buf[0].device != flat.device:
            buf = [torch.zeros(flat.numel(), dtype=torch.float32, device=flat.device), 0]
        err, step = buf
        step += 1
        reset = step >= reset_T
        q, qp = Q.loco_quantize(flat.contiguous(), err, (per // gs) * world, num_bits, beta, reset)
        try:
            p.intra_ef_buf = [err, 0 if reset else step]
        except Exception:
            pass
        outs.append(_exchange_and_reduce(q, qp, world, rank, per0, per, gs, n, num_bits, g.dtype, local))
    return outs
