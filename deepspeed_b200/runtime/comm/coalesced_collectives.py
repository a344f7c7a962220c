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
        per0 = math.ceil(n / world)              # elements this rank ends up owning
        gs = Q.aligned_group_size(per0)          # device kernels: groups of a multiple of 8 elements, never across ranks
        per = (per0 + gs - 1) // gs * gs         # per-rank chunk padded to whole groups
        if per * world != n:
            full = flat.new_zeros(per * world)
            for r in range(world):               # rank r's slice [r * per0, (r + 1) * per0) moves to chunk r
                lo_r, hi_r = r * per0, min(n, (r + 1) * per0)
                if hi_r > lo_r:
                    full[r * per:r * per + hi_r - lo_r].copy_(flat[lo_r:hi_r])
            flat = full
        per_rank_groups = per // gs
        q, params = Q.quantize(flat.contiguous(), per_rank_groups * world, num_bits, Q.Symmetric)
        q_recv, p_recv = torch.empty_like(q), torch.empty_like(params)
        dist.all_to_all_single(q_recv.view(-1), q.view(-1), group=local)
        dist.all_to_all_single(p_recv.view(-1), params.view(-1), group=local)
        deq = Q.dequantize(q_recv, p_recv, per_rank_groups * world, num_bits, Q.Symmetric, dtype=torch.float32)
        red = deq.view(world, per).sum(0).div_(world)
        lo = rank * per0
        valid = max(0, min(per0, n - lo))
        out.append(red[:valid].to(t.dtype))
    return out


def all_to_all_loco_quant_reduce(params, groups: dict = None, loco_param: dict = None, num_bits=4):
    """LoCo-ZeRO++: qgZ with an error-feedback buffer per tensor (``p.intra_ef_buf``) so quantisation error is
    re-injected at the next step (reference :30)."""
    loco_param = loco_param or {}
    beta = float(loco_param.get("err_beta", 0.8))
    reset_T = int(loco_param.get("reset_T", 1024))
    outs = []
    for p in params:
        g = p.grad if hasattr(p, "grad") and p.grad is not None else p
        buf = getattr(p, "intra_ef_buf", None)
        if buf is None or buf[0].shape != g.shape:
            buf = [torch.zeros_like(g, dtype=torch.float32), 0]
        err, step = buf
        comp = g.float() + err
        qd = Q.fake_quantize(comp.reshape(-1).contiguous(), _groups_for(comp), num_bits, Q.Symmetric).view_as(comp)
        new_err = comp - qd
        step += 1
        if step >= reset_T:
            err.zero_()
            step = 0
        else:
            err.mul_(beta).add_(new_err, alpha=1 - beta)
        try:
            p.intra_ef_buf = [err, step]
        except Exception:
            pass
        outs.extend(all_to_all_quant_reduce([comp.to(g.dtype)], groups, num_bits))
    return outs
