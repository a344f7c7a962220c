"""Stage-to-stage activation / gradient transport (reference ``runtime/pipe/p2p.py:46,67`` + the tensor-meta
handshake of ``pipe/engine.py:928-1010``).  A message is a tensor or a tuple of tensors; the first message of
each (src, dst, kind) stream is preceded by a small int64 header describing shapes / dtypes so the receiver can
allocate, afterwards shapes are assumed static (``dynamic_shape=True`` re-sends the header every time)."""
import torch

from deepspeed_b200 import comm as dist

_DTYPES = [torch.float32, torch.float16, torch.bfloat16, torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8,
           torch.bool, torch.float64]
_grid = None
_meta_cache = {}


def init_process_groups(grid):
    global _grid
    _grid = grid
    _meta_cache.clear()


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if (torch.cuda.is_available() and dist.get_backend() == "nccl") \
        else torch.device("cpu")


def _encode(tensors):
    h = [len(tensors)]
    for t in tensors:
        h += [_DTYPES.index(t.dtype), int(t.requires_grad), t.dim()] + list(t.shape)
    return torch.tensor(h + [0] * (64 - len(h)), dtype=torch.int64, device=_dev())


def _decode(h):
    h = h.tolist()
    n, pos, out = h[0], 1, []
    for _ in range(n):
        dt, rg, nd = h[pos:pos + 3]
        shape = h[pos + 3:pos + 3 + nd]
        pos += 3 + nd
        out.append((_DTYPES[dt], bool(rg), tuple(shape)))
    return out


def send_obj(obj, dst_stage, kind, dynamic=False):
    """Send a tensor / tuple of tensors to pipeline stage ``dst_stage``."""
    dst = _grid.stage_to_global(stage_id=dst_stage)
    tensors = (obj, ) if torch.is_tensor(obj) else tuple(obj)
    key = ("s", dst, kind)
    if dynamic or key not in _meta_cache:
        dist.send(_encode(tensors), dst)
        _meta_cache[key] = True
    for t in tensors:
        t = t.detach().contiguous()
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        dist.send(t.to(_dev()) if t.device != _dev() else t, dst)


def recv_obj(src_stage, kind, dynamic=False):
    src = _grid.stage_to_global(stage_id=src_stage)
    key = ("r", src, kind)
    if dynamic or key not in _meta_cache:
        h = torch.zeros(64, dtype=torch.int64, device=_dev())
        dist.recv(h, src)
        _meta_cache[key] = _decode(h)
    out = []
    for dt, rg, shape in _meta_cache[key]:
        buf = torch.empty(shape, dtype=torch.uint8 if dt == torch.bool else dt, device=_dev())
        dist.recv(buf, src)
        if dt == torch.bool:
            buf = buf.bool()
        if rg and buf.is_floating_point():
            buf.requires_grad_(True)
        out.append(buf)
    return out[0] if len(out) == 1 else tuple(out)


def send_recv(send_obj_, peer_stage, send_kind, recv_kind, dynamic=False):
    """Exchange with ONE neighbour in a single grouped operation: post the sends and the receives together
    (``batch_isend_irecv``) so two stages that both need to send first cannot dead-lock.  Returns the
    received object."""
    peer = _grid.stage_to_global(stage_id=peer_stage)
    tensors = (send_obj_, ) if torch.is_tensor(send_obj_) else tuple(send_obj_)
    skey, rkey = ("s", peer, send_kind), ("r", peer, recv_kind)
    need_sh = dynamic or skey not in _meta_cache
    need_rh = dynamic or rkey not in _meta_cache
    if need_sh or need_rh:
        ops = []
        if need_sh:
            ops.append(dist.P2POp(torch.distributed.isend, _encode(tensors), peer))
        hbuf = torch.zeros(64, dtype=torch.int64, device=_dev())
        if need_rh:
            ops.append(dist.P2POp(torch.distributed.irecv, hbuf, peer))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        if need_sh:
            _meta_cache[skey] = True
        if need_rh:
            _meta_cache[rkey] = _decode(hbuf)
    ops, keep = [], []
    for t in tensors:
        t = t.detach().contiguous()
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        t = t.to(_dev()) if t.device != _dev() else t
        keep.append(t)
        ops.append(dist.P2POp(torch.distributed.isend, t, peer))
    bufs = []
    for dt, rg, shape in _meta_cache[rkey]:
        buf = torch.empty(shape, dtype=torch.uint8 if dt == torch.bool else dt, device=_dev())
        bufs.append((buf, dt, rg))
        ops.append(dist.P2POp(torch.distributed.irecv, buf, peer))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    out = []
    for buf, dt, rg in bufs:
        if dt == torch.bool:
            buf = buf.bool()
        if rg and buf.is_floating_point():
            buf.requires_grad_(True)
        out.append(buf)
    return out[0] if len(out) == 1 else tuple(out)


# --- raw tensor send/recv between adjacent stages (reference ``runtime/pipe/p2p.py:46-93``) --------------------------
_pending = []


def can_send_recv() -> bool:
    """True when the backend has real point-to-point primitives (always, for NCCL ≥ 2.7 and gloo)."""
    return True


def _check_adjacent(src_stage, dest_stage):
    n = _grid.pipe_parallel_size
    first, last = 0, n - 1
    ok = abs(src_stage - dest_stage) == 1 or {src_stage, dest_stage} == {first, last}
    assert ok, f"Functionality currently limited to send and receive between adjacent ranks only ({src_stage}->{dest_stage})"


def send(tensor, dest_stage, async_op=False):
    """Send ``tensor`` (pre-agreed shape) to the adjacent pipeline stage ``dest_stage``."""
    _check_adjacent(_grid.get_stage_id(), dest_stage)
    dst = _grid.stage_to_global(stage_id=dest_stage)
    if async_op:
        _pending.append(dist.isend(tensor, dst))
        return _pending[-1]
    return dist.send(tensor, dst)


def recv(tensor, src_stage, async_op=False):
    """Receive into ``tensor`` from the adjacent pipeline stage ``src_stage``."""
    _check_adjacent(src_stage, _grid.get_stage_id())
    src = _grid.stage_to_global(stage_id=src_stage)
    if async_op:
        _pending.append(dist.irecv(tensor, src))
        return _pending[-1]
    return dist.recv(tensor, src)


def wait():
    """Complete every outstanding async ``send`` / ``recv``."""
    while _pending:
        _pending.pop(0).wait()
    if torch.cuda.is_available() and dist.get_backend() == "nccl":
        torch.cuda.current_stream().synchronize()
