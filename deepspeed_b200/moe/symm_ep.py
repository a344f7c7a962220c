"""Expert-parallel token exchange fused with the permutation kernels, over symmetric (NVLink peer-mapped) memory.

``dispatch``  = scatter ⊕ all-to-all:  every token row is stored directly into the receive buffer of the rank owning
its expert, already in the ``[E_local, ep*C, H]`` layout the experts consume.
``combine``   = all-to-all ⊕ gather:   every token reads its k expert outputs straight out of the owners' buffers and
mixes them with the gate weights.
Both are autograd functions whose backward uses the same two kernels with roles swapped.  Replaces the
``scatter → all_to_all_single → … → all_to_all_single → gather`` pipeline of the NCCL path
(reference ``moe/sharded_moe.py:97 _AllToAll``, ``:609``, ``:669``); device code in ``csrc/cuda/moe_symm.cu``.

Two symmetric buffers per EP group: ``buf_in`` (peers *write* into it) and ``buf_out`` (peers *read* from it).  Ranks are
ordered by the barrier kernels of ``comm/symm_impl.py``; results are cloned out of ``buf_in`` so the next layer can reuse it.
"""
import ctypes

import torch

from deepspeed_b200 import comm as dist
from deepspeed_b200.ops import native as N


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class SymmEP:
    """Per-EP-group state: symmetric context + the two exchange buffers."""
    _cache = {}

    @classmethod
    def get(cls, group):
        from deepspeed_b200.comm import symm
        key = id(group)
        if key not in cls._cache:
            ok = symm.is_supported(group, explicit=False)
            cls._cache[key] = cls(symm.get_context(group), group) if ok else None
        return cls._cache[key]

    def __init__(self, ctx, group):
        self.ctx = ctx
        self.group = group
        self.ep = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.buf_in = self.buf_out = None
        self.numel = 0
        self.dtype = None

    def ensure(self, rows, hidden, dtype):
        n = rows * hidden
        if self.buf_in is None or n > self.numel or dtype != self.dtype:
            # collective allocation: every rank reaches this with identical shapes (same capacity, same hidden)
            self.buf_in = self.ctx.alloc(n, dtype)
            self.buf_out = self.ctx.alloc(n, dtype)
            self.numel, self.dtype = n, dtype
        return self.buf_in[:n], self.buf_out[:n]

    def peers(self, t):
        arr, _ = self.ctx._peers(t)
        return arr

    # -- the two device primitives -------------------------------------------------------------------------------
    def scatter_to_peers(self, x, dst, expert_ids, positions, K, C, e_local, weights=None, eo=None, dweights=None):
        rc = N.cuda().dsb_moe_scatter_peer(_p(x), self.peers(dst), _p(expert_ids), _p(positions), _p(weights),
                                           self.peers(eo) if eo is not None else ctypes.c_void_p(0), _p(dweights),
                                           expert_ids.numel(), K, x.shape[-1], C, e_local, self.ep, self.rank, N.dt(x),
                                           N.stream())
        N.check(rc, "moe_scatter_peer")

    def gather_from_peers(self, src, y, expert_ids, positions, K, C, e_local, weights=None):
        rc = N.cuda().dsb_moe_gather_peer(self.peers(src), _p(y), _p(expert_ids), _p(positions), _p(weights), y.shape[0], K,
                                          y.shape[-1], C, e_local, self.ep, self.rank, N.dt(y), N.stream())
        N.check(rc, "moe_gather_peer")


class _Dispatch(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, expert_ids, positions, st, K, C, e_local):
        rows = e_local * st.ep * C
        H = x.shape[-1]
        buf_in, _ = st.ensure(rows, H, x.dtype)
        buf_in.zero_()                       # rows nobody writes must be exact zeros (they enter the expert GEMMs)
        st.ctx.barrier()                     # everyone zeroed / finished reading the previous contents
        st.scatter_to_peers(x.contiguous(), buf_in, expert_ids, positions, K, C, e_local)
        st.ctx.barrier()                     # all peers' rows have landed
        ctx.st, ctx.K, ctx.C, ctx.e_local, ctx.T = st, K, C, e_local, x.shape[0]
        ctx.save_for_backward(expert_ids, positions)
        return buf_in.view(e_local, st.ep * C, H).clone()

    @staticmethod
    def backward(ctx, d_recv):
        expert_ids, positions = ctx.saved_tensors
        st = ctx.st
        H = d_recv.shape[-1]
        _, buf_out = st.ensure(ctx.e_local * st.ep * ctx.C, H, d_recv.dtype)
        buf_out.copy_(d_recv.reshape(-1))
        st.ctx.barrier()
        dx = torch.empty(ctx.T, H, dtype=d_recv.dtype, device=d_recv.device)
        st.gather_from_peers(buf_out, dx, expert_ids, positions, ctx.K, ctx.C, ctx.e_local)
        st.ctx.barrier()                     # peers are done reading buf_out before it is overwritten
        return dx, None, None, None, None, None, None


class _Combine(torch.autograd.Function):

    @staticmethod
    def forward(ctx, eo, weights, expert_ids, positions, st, K, C, e_local, T):
        H = eo.shape[-1]
        _, buf_out = st.ensure(e_local * st.ep * C, H, eo.dtype)
        buf_out.copy_(eo.reshape(-1))
        st.ctx.barrier()
        y = torch.empty(T, H, dtype=eo.dtype, device=eo.device)
        w = weights.reshape(-1).to(torch.float32).contiguous()
        st.gather_from_peers(buf_out, y, expert_ids, positions, K, C, e_local, weights=w)
        st.ctx.barrier()
        ctx.st, ctx.K, ctx.C, ctx.e_local = st, K, C, e_local
        ctx.w_shape, ctx.w_dtype = weights.shape, weights.dtype
        ctx.save_for_backward(eo, w, expert_ids, positions)
        return y

    @staticmethod
    def backward(ctx, dy):
        eo, w, expert_ids, positions = ctx.saved_tensors
        st = ctx.st
        H = eo.shape[-1]
        rows = ctx.e_local * st.ep * ctx.C
        buf_in, buf_out = st.ensure(rows, H, eo.dtype)
        buf_out.copy_(eo.reshape(-1))        # republish the expert outputs: peers dot them with their dy rows
        buf_in.zero_()
        st.ctx.barrier()
        dw = torch.empty(w.numel(), dtype=torch.float32, device=dy.device)
        st.scatter_to_peers(dy.contiguous(), buf_in, expert_ids, positions, ctx.K, ctx.C, ctx.e_local, weights=w, eo=buf_out,
                            dweights=dw)
        st.ctx.barrier()
        d_eo = buf_in.view(eo.shape).clone()
        return d_eo, dw.view(ctx.w_shape).to(ctx.w_dtype), None, None, None, None, None, None, None


def dispatch(st: SymmEP, x, expert_ids, positions, K, C, e_local):
    """x [T, H] -> [E_local, ep*C, H] (rows of source rank s, expert e at [e, s*C:(s+1)*C])."""
    return _Dispatch.apply(x, expert_ids.reshape(-1).to(torch.int32).contiguous(), positions.to(torch.int32).contiguous(),
                           st, K, C, e_local)


def combine(st: SymmEP, eo, weights, expert_ids, positions, K, C, e_local, T):
    """eo [E_local, ep*C, H] -> y [T, H] = sum_k w[t,k] * eo_owner[row(t,k)]."""
    return _Combine.apply(eo.contiguous(), weights, expert_ids.reshape(-1).to(torch.int32).contiguous(),
                          positions.to(torch.int32).contiguous(), st, K, C, e_local, T)
